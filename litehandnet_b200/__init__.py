"""litehandnet_b200 — B200-native heatmap encode/decode hot path."""
