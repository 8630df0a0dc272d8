"""Mirror of the reference's ``utils`` package for the hot path (tower_extraction, elevation_converter)."""
