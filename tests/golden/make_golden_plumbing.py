"""Golden fixture for BASELINE configs[0] (``configs/test_small.yml`` plumbing), made by RUNNING THE REFERENCE.

    python tests/golden/make_golden_plumbing.py        # build container only (needs /root/reference)

What it does, with the reference's own code only (SURVEY.md appendix A route (ii): the full ``twotower``
package is importable once ``tools.huggingface`` is stubbed):

  random.seed(0) -> dataset_factory.generate_synthetic_pairs(200, 2) -> transform_and_save_dataset(.., 'triplets')
  -> load_config('configs/test_small.yml') -> build_pipeline(config, 'cpu')   (twotower/train.py:298-371)
  -> DataLoader(batch_size=32, shuffle=False) -> train_epoch(...)             (twotower/train.py:64-160)

and records the triplet texts, the initial state_dict, every batch's loss (captured by wrapping ``loss_fn``) and the
final state_dict.  ``tests/test_plumbing_golden.py`` replays the same loop shape with the ``*_b200`` classes on
the GPU box (where ``/root/reference`` does not exist) and compares the per-batch loss trajectory.
"""
from __future__ import annotations

import json
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get("TT_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    os.environ.setdefault("WANDB_MODE", "disabled")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import tools                                           # namespace package of the reference
    stub = types.ModuleType("tools.huggingface")
    stub.setup_repository = lambda *a, **k: None
    stub.upload_model_to_hub = lambda *a, **k: None
    stub.save_and_upload = lambda *a, **k: None
    sys.modules["tools.huggingface"] = stub
    import twotower                                        # noqa: F401
    import twotower.train as ref_train
    import twotower.utils as ref_utils
    return ref_train, ref_utils


def main():
    ref_train, ref_utils = import_reference()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)                                      # dataset_factory writes ./data/{raw,processed}
        try:
            from dataset_factory.synthetic_generators import generate_synthetic_pairs
            from dataset_factory.utils import transform_and_save_dataset
            random.seed(0)
            tsv = generate_synthetic_pairs(n_positive=200, n_negative_per_positive=2, output_file="pairs.tsv")
            pq = transform_and_save_dataset(tsv, "plumbing_triplets.parquet", "triplets")
            config = ref_utils.load_config(os.path.join(REF, "configs", "test_small.yml"))
            config.update(data=str(pq), use_wandb=False, device="cpu")
            config["huggingface"] = {"push_to_hub": False}
            torch.manual_seed(0)
            model, dataset, optimizer, loss_fn = ref_train.build_pipeline(config, "cpu")
            init = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
            losses = []

            def recording_loss(q, p, n):
                out = loss_fn(q, p, n)
                losses.append(float(out.detach()))
                return out
            loader = torch.utils.data.DataLoader(dataset, batch_size=config["batch_size"], shuffle=False)
            metrics = ref_train.train_epoch(model, loader, optimizer, recording_loss, "cpu", use_wandb=False)
            final = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
            meta = {
                "config": {k: config[k] for k in ("tokeniser", "embedding", "encoder", "loss", "optimizer", "batch_size")},
                "vocab": dataset.tokeniser.string_to_index,
                "n_triplets": len(dataset),
                "queries": dataset.query_texts, "positives": dataset.positive_doc_texts,
                "negatives": dataset.negative_doc_texts,
                "first_encoded": [dataset.encoded_queries[0], dataset.encoded_positive_docs[0], dataset.encoded_negative_docs[0]],
                "batch_losses": losses, "epoch_loss": float(metrics["loss"]),
                "torch": torch.__version__,
            }
        finally:
            os.chdir(cwd)
    with open(os.path.join(OUT, "plumbing_test_small.json"), "w") as f:
        json.dump(meta, f)
    np.savez_compressed(os.path.join(OUT, "plumbing_test_small.npz"),
                        **{f"init/{k}": v for k, v in init.items()}, **{f"final/{k}": v for k, v in final.items()})
    print(f"{len(losses)} batches, loss {losses[0]:.6f} -> {losses[-1]:.6f}, epoch loss {metrics['loss']:.6f}, "
          f"{meta['n_triplets']} triplets, vocab {len(meta['vocab'])}")


if __name__ == "__main__":
    main()
