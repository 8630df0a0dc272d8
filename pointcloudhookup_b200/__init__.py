"""pointcloudhookup_b200 — B200-native (sm_100a) implementation of pointcloudhookup's per-point
LAS hot path behind the reference's Python call surface.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"


def install_dropin() -> None:
    """Register this package's mirrors under the reference's import names (``ui.import_PC``,
    ``ui.Sampling``, ``ui.extract``, ``ui.compress``, ``utils.tower_extraction``,
    ``utils.elevation_converter``, ``crs``) so pyGUI_towers_test.py's imports resolve to the B200
    path.  See INTEGRATION.md."""
    import importlib
    import sys
    import types
    for pkg in ("ui", "utils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    for name in ("ui.import_PC", "ui.Sampling", "ui.extract", "ui.compress", "utils.tower_extraction",
                 "utils.elevation_converter", "crs"):
        mod = importlib.import_module(f"{__name__}.{name}")
        sys.modules[name] = mod
        if "." in name:
            setattr(sys.modules[name.split(".")[0]], name.split(".")[1], mod)
