"""`ps.metrics` functions that consume the radius map, evaluated on its index form:
`pore_size_distribution` (/root/reference/src/porespy/metrics/_funcs.py:558-632) and the `sizes` branch of
`pc_curve` (/root/reference/src/porespy/metrics/_funcs.py:980-1090).  See porespy_b200/sizemap.py."""
from .sizemap import pc_curve, pore_size_distribution

__all__ = ["pore_size_distribution", "pc_curve"]
