"""B200-native detect+track hot path."""
