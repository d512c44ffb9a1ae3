"""smrt_b200 — B200-native DORT hot path for SMRT."""
