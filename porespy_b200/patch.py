"""Monkey-patch points for unmodified PoreSpy scripts:
`sys.modules['edt']`, `porespy.filters.{porosimetry, local_thickness, trim_disconnected_blobs}` (and the
other flood users: `find_disconnected_voxels`, `fill_blind_pores`, `trim_floating_solid`,
`trim_nonpercolating_paths`)
and the `edt` name bound inside `porespy.filters._funcs` / `porespy.tools._funcs`."""
import sys
import types

_saved = {}


def install(patch_edt_module=True):
    """Route PoreSpy's hot path through porespy_b200.  Safe to call before or after
    `import porespy`."""
    from . import edt as edt_mod
    from . import filters as f
    if patch_edt_module and "edt" not in _saved:
        _saved["edt"] = sys.modules.get("edt")
        shim = types.ModuleType("edt")
        shim.edt, shim.edtsq = edt_mod.edt, edt_mod.edtsq
        shim.__doc__ = "porespy_b200 drop-in for the `edt` package"
        sys.modules["edt"] = shim
    ps = sys.modules.get("porespy")
    if ps is not None and "porespy" not in _saved:
        _saved["porespy"] = {}
        for name in ("porosimetry", "local_thickness", "trim_disconnected_blobs", "find_disconnected_voxels",
                     "fill_blind_pores", "trim_floating_solid", "trim_nonpercolating_paths", "find_trapped_regions"):
            for mod in (ps.filters, getattr(ps.filters, "_funcs", None)):
                if mod is not None and hasattr(mod, name):
                    _saved["porespy"][(mod, name)] = getattr(mod, name)
                    setattr(mod, name, getattr(f, name))
        for modname in ("porespy.filters._funcs", "porespy.tools._funcs"):
            mod = sys.modules.get(modname)
            if mod is not None and hasattr(mod, "edt"):
                _saved["porespy"][(mod, "edt")] = mod.edt
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
