"""Drop-in registration into an importable copy of the reference.

``install_into_reference()`` adds the B200 classes to the reference's own registries under
new names, so an unmodified reference config selects them with e.g.
``embedding.type: lookup_b200`` / ``encoder.arch: mean_b200`` / ``loss.type: in_batch_b200``.
Nothing here runs unless the reference package is importable (it is not on the GPU box).
"""
from __future__ import annotations

import importlib
from typing import Dict


def install_into_reference(twotower_pkg: str = "twotower", search_pkg: str = "inference.search") -> Dict[str, list]:
    from . import embeddings, encoders, losses, search, tokenisers

    done: Dict[str, list] = {}
    ref_emb = importlib.import_module(f"{twotower_pkg}.embeddings")
    ref_emb.REGISTRY.update({"lookup_b200": embeddings.LookupEmbedding})
    done["embeddings"] = ["lookup_b200"]
    ref_enc = importlib.import_module(f"{twotower_pkg}.encoders")
    ref_enc.TOWER_REGISTRY.update({"mean_b200": encoders.MeanPoolingTower,
                                   "avg_pool_b200": encoders.AveragePoolingTower})
    done["encoders"] = ["mean_b200", "avg_pool_b200"]
    ref_loss = importlib.import_module(f"{twotower_pkg}.losses")
    ref_loss.LOSS_REGISTRY.update({"triplet_b200": losses.contrastive_triplet_loss,
                                   "multiple_negatives_b200": losses.multiple_negatives_loss,
                                   "in_batch_b200": losses._in_batch_adapter})
    done["losses"] = ["triplet_b200", "multiple_negatives_b200", "in_batch_b200"]
    ref_tok = importlib.import_module(f"{twotower_pkg}.tokenisers")
    ref_tok.REGISTRY.update({"char_b200": tokenisers.CharTokeniser, "word_b200": tokenisers.WordTokeniser})
    done["tokenisers"] = ["char_b200", "word_b200"]
    try:
        ref_search = importlib.import_module(search_pkg)
        setattr(ref_search, "B200TwoTowerSearch", search.TwoTowerSearch)
        done["search"] = ["B200TwoTowerSearch"]
    except Exception:                                           # search package optional
        done["search"] = []
    return done
