"""B200-native tracker hot path (ROI Align + association step)."""
