"""two_towers_b200 -- B200-native (sm_100a) two-tower training + retrieval hot path.

A from-scratch implementation of the hot path of k0r1g/two-towers behind the reference's own
plugin API: ``tokenisers.REGISTRY``, ``embeddings.REGISTRY``, ``encoders.TOWER_REGISTRY``,
``losses.LOSS_REGISTRY`` and ``search.BaseSearch / TwoTowerSearch``.  All arithmetic runs in
hand-written CUDA kernels inside ``csrc/libtt_b200.so`` (C ABI: ``include/tt_b200.h``); there
is no CPU / eager-PyTorch fallback.
"""
from . import _lib, dataset, embeddings, encoders, evaluate, losses, ops, parallel, search, tokenisers, train  # noqa: F401
from .dataset import TripletDataset
from .embeddings import BaseEmbedding, LookupEmbedding, PretrainedEmbedding
from .encoders import (AveragePoolingTower, BaseTower, MeanPoolingTower, TwoTower, TOWER_REGISTRY, build_tower,
                       build_two_tower)
from .integration import install_into_reference
from .losses import (LOSS_REGISTRY, contrastive_triplet_loss, in_batch_sampled_softmax_loss,
                     multiple_negatives_loss)
from .ops import get_default_precision, set_default_precision
from .search import BaseSearch, TwoTowerSearch
from .tokenisers import BaseTokeniser, CharTokeniser, WordTokeniser
from .train import FusedTrainer

__version__ = "0.1.0"
