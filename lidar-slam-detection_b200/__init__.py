"""lsdreg — B200-native registration hot path for LSD (lidar-slam-detection).  See DESIGN.md."""
