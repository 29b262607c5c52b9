import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import two_towers_b200 as tt
import bench
dev = torch.device("cuda", 0)
c = bench.MSM
def run(par):
    os.environ["TT_TOWER_PAR"] = par
    torch.manual_seed(0)
    emb = tt.embeddings.build("lookup", c["V"], embedding_dim=c["E"])
    model = tt.build_two_tower("mean", emb, hidden_dim=c["H"], tied_weights=False).to(dev)
    q, d = (bench.repeated_ids(c["B"], c["L"], c["V"], s_, c["repeat"]).to(dev) for s_ in (1, 2))
    tr = tt.FusedTrainer(model, loss="in_batch", batch_size=c["B"], max_len=c["L"], precision="bf16", use_cuda_graph=True, id_dtype=torch.int32)
    assert tr.par_towers == (par == "1"), tr.par_towers
    losses = [tr.step(q, d).item() for _ in range(4)]
    torch.cuda.synchronize()
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): tr.step(q, d)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 10)
    return losses, float(np.median(ts)), tr.flat.clone(), tr.kernels_per_step()
l1, t1, f1, k1 = run("1")
l0, t0, f0, k0 = run("0")
print("parallel towers: %.1f us/step (%d launches)   serial: %.1f us/step (%d launches)" % (t1, k1, t0, k0))
print("losses", l1, l0)
print("max param diff after the same number of steps", (f1 - f0).abs().max().item())
