#!/usr/bin/env python
"""Golden vectors for the ranking metrics: runs the REFERENCE's twotower/evaluate.py helpers in the build
container on seeded relevance lists and writes tests/golden/eval_metrics.json.

    python tests/golden/make_golden_eval.py        # needs /root/reference (not available on the GPU box)
"""
import importlib.util
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_eval", "/root/reference/twotower/evaluate.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(2024)
    cases = []
    for i in range(120):
        n = int(rng.integers(1, 20))
        rel = rng.integers(0, 2, n).tolist()
        if i % 10 == 0:
            rel = [0] * n                                   # no relevant document
        k = int(rng.integers(1, 15))
        if max(n, k) == 1:
            k = 2                                           # sklearn's ndcg_score rejects a single document
        cases.append({"relevance": rel, "k": k,
                      "mrr": float(ref.mean_reciprocal_rank(rel)),
                      "precision": float(ref.precision_at_k(rel, k)),
                      "recall": float(ref.recall_at_k(rel, k, int(np.sum(rel)))),
                      "ndcg": float(ref.ndcg_at_k(rel, k))})
    with open(os.path.join(HERE, "eval_metrics.json"), "w") as f:
        json.dump(cases, f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
