"""credgcn: B200-native credibility-aware LightGCN hot path."""
__version__ = "0.1.0"
