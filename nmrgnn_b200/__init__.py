"""nmrgnn_b200 — B200-native (sm_100a) forward path for the nmrgnn chemical-shift GNN."""
__version__ = "0.1.0"
