"""credgcn: B200-native credibility-aware LightGCN hot path.

Importing the package never touches CUDA; the first op loads libcredgcn.so and raises if the
sm_100a extension is missing or a tensor is not on a CUDA device (no CPU fallback)."""
__version__ = "0.1.0"
