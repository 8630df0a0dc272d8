"""Mirror of the reference's ``ui`` package for the hot path (import_PC, Sampling, extract, compress)."""
