"""B200-native (sm_100a) embedding -> feature-interaction -> sparse-update hot path, behind the nn.Module /
Trainer / optimizer protocol of WardellZc/DeepLearningRecommendationSystem.

Layout mirrors the reference: ``model/`` (drop-in modules), ``trainer/``, ``sampler/``; ``nfield`` holds the
N-field generalisations used by the synthetic Criteo-shaped configs; ``csrc/`` + ``include/recsys_b200.h`` are the
CUDA kernels and their C ABI, bound in ``_lib`` / ``ops``.
"""
__version__ = "0.1.0"
