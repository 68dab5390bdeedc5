"""B200-native GATRes hot path."""
