"""Per-parameter gradient error of one bf16 trainer step against the fp32 step (triplet and in-batch loss, B = 384)."""
import copy, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import two_towers_b200 as tt
DEV='cuda'
for loss in ("triplet", "in_batch"):
    torch.manual_seed(1)
    emb = tt.embeddings.build("lookup", 128, embedding_dim=64)
    m32 = tt.build_two_tower("mean", emb, hidden_dim=256, tied_weights=True).to(DEV)
    m16 = copy.deepcopy(m32)
    g = torch.Generator().manual_seed(4)
    B, L = 384, 64
    q, d, n = (torch.randint(0, 128, (B, L), generator=g) for _ in range(3))
    t32 = tt.FusedTrainer(m32, loss=loss, batch_size=B, max_len=L, precision="fp32", use_cuda_graph=False)
    t16 = tt.FusedTrainer(m16, loss=loss, batch_size=B, max_len=L, precision="bf16", use_cuda_graph=False)
    args = (q, d, n) if loss == "triplet" else (q, d)
    t32.step(*args); t16.step(*args)
    print(loss, "loss", t32.loss.item(), t16.loss.item())
    for (name, p32), (_, p16) in zip(m32.named_parameters(), m16.named_parameters()):
        a, b = p16.grad.double(), p32.grad.double()
        print(f"  {name:50s} max|g32| {b.abs().max():.3e}  max err {(a-b).abs().max():.3e}  rel {(a-b).abs().max()/b.abs().max():.3e}  rel-l2 {((a-b).norm()/b.norm()):.3e}")
