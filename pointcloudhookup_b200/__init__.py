"""pointcloudhookup_b200 — B200-native (sm_100a) implementation of pointcloudhookup's per-point
LAS hot path behind the reference's Python call surface.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
