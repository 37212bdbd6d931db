"""Monkey-patch points for unmodified PoreSpy scripts:
`sys.modules['edt']`, `porespy.filters.{porosimetry, local_thickness, trim_disconnected_blobs}` (and the
other flood users: `find_disconnected_voxels`, `fill_blind_pores`, `trim_floating_solid`,
`trim_nonpercolating_paths`, `find_trapped_regions`), and the `edt` name that every PoreSpy module bound with
`from edt import edt` (46 call sites, SURVEY 8(f) rank 1: filters/_funcs.py:5, tools/_funcs.py:6,
filters/_snows.py, simulations/_drainage.py, simulations/_ibip.py, networks/_getnet.py,
metrics/_regionprops.py, generators/_imgen.py, generators/_pseudo_packings.py, io/_funcs.py, beta/*)."""
import sys
import types

_saved = {}

_FILTERS = ("porosimetry", "local_thickness", "trim_disconnected_blobs", "find_disconnected_voxels",
            "fill_blind_pores", "trim_floating_solid", "trim_nonpercolating_paths", "find_trapped_regions",
            "size_to_seq", "size_to_satn", "seq_to_satn")


def install(patch_edt_module=True):
    """Route PoreSpy's hot path through porespy_b200.  Safe to call before or after `import porespy`:
    before, the `edt` shim in `sys.modules` is what PoreSpy's `from edt import edt` lines pick up; after,
    every already imported `porespy.*` module whose `edt` attribute is the original function is rebound."""
    import importlib
    edt_mod = importlib.import_module(".edt", __package__)     # (the package attribute `edt` is the function)
    f = importlib.import_module(".filters", __package__)
    if patch_edt_module and "edt" not in _saved:
        old = sys.modules.get("edt")
        _saved["edt"] = old
        shim = types.ModuleType("edt")
        shim.edt, shim.edtsq = edt_mod.edt, edt_mod.edtsq
        shim.__doc__ = "porespy_b200 drop-in for the `edt` package"
        sys.modules["edt"] = shim
    ps = sys.modules.get("porespy")
    if ps is not None and "porespy" not in _saved:
        saved = _saved["porespy"] = {}
        for name in _FILTERS:
            for mod in (getattr(ps, "filters", None), sys.modules.get("porespy.filters._funcs"),
                        sys.modules.get("porespy.filters._size_seq_satn")):
                if mod is not None and hasattr(mod, name):
                    saved[(mod, name)] = getattr(mod, name)
                    setattr(mod, name, getattr(f, name))
        from . import metrics as m
        for mod in (getattr(ps, "metrics", None), sys.modules.get("porespy.metrics._funcs")):
            if mod is not None and hasattr(mod, "pore_size_distribution"):
                saved[(mod, "pore_size_distribution")] = mod.pore_size_distribution
                mod.pore_size_distribution = m.pore_size_distribution
        old_edt = _saved.get("edt")
        originals = {getattr(old_edt, "edt", None)} - {None}
        for modname, mod in list(sys.modules.items()):
            if mod is None or not (modname == "porespy" or modname.startswith("porespy.")):
                continue
            cur = mod.__dict__.get("edt")
            if cur is None or isinstance(cur, types.ModuleType) or cur is edt_mod.edt:
                continue
            # `from edt import edt` binds the function; rebind it when it is the original package's
            # function (or, when the original package is unknown, any callable of that name)
            if callable(cur) and (not originals or cur in originals):
                saved[(mod, "edt")] = cur
                mod.edt = edt_mod.edt


def uninstall():
    if "edt" in _saved:
        old = _saved.pop("edt")
        if old is None:
            sys.modules.pop("edt", None)
        else:
            sys.modules["edt"] = old
    for (mod, name), val in _saved.pop("porespy", {}).items():
        setattr(mod, name, val)
