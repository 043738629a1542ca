"""CPU: every `cu.*` / `models.*` call that class SimuBasin of the reference's wmf.py makes exists on the drop-in objects and
accepts that many positional arguments (the reference calls them positionally, dims included).  Parsed from the reference's
source when the tree is present (it is where the driver runs the CPU suite); nothing is executed."""
import ast
import inspect
import os
import warnings

import pytest

WMF_PY = "/root/reference/wmf/wmf.py"
# call sites that are optional paths outside the hot path
ALLOWED_MISSING = set()


@pytest.mark.skipif(not os.path.exists(WMF_PY), reason="reference tree not present")
def test_simubasin_call_sites_are_covered():
    from wmf_b200 import CuModule, ModelsModule
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tree = ast.parse(open(WMF_PY, encoding="latin-1").read())
    objs = {"cu": CuModule, "models": ModelsModule}
    calls = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "SimuBasin":
            for c in ast.walk(node):
                if (isinstance(c, ast.Call) and isinstance(c.func, ast.Attribute) and isinstance(c.func.value, ast.Name)
                        and c.func.value.id in objs):
                    calls.add((c.func.value.id, c.func.attr, len(c.args), c.lineno))
    assert len(calls) > 20
    problems = []
    for mod, attr, nargs, ln in sorted(calls, key=lambda t: t[-1]):
        m = getattr(objs[mod], attr, None)
        if m is None:
            if attr not in ALLOWED_MISSING:
                problems.append(f"wmf.py:{ln} {mod}.{attr} is missing")
            continue
        params = [p for p in inspect.signature(m).parameters.values() if p.name != "self"]
        pos = [p for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        need = len([p for p in pos if p.default is p.empty])
        if not need <= nargs <= len(pos):
            problems.append(f"wmf.py:{ln} {mod}.{attr} called with {nargs} positional arguments, the shim takes {need}..{len(pos)}")
    assert not problems, "\n".join(problems)


@pytest.mark.skipif(not os.path.exists(WMF_PY), reason="reference tree not present")
def test_methods_reached_from_the_model_workflow_are_covered():
    """Transitive closure: every method of Basin / SimuBasin reachable through ``self.*`` calls from the modelling workflow
    (construction, geomorphology, control points, storages, sediments / slides / floods set-up, rain interpolation, run_shia)
    only uses ``cu.*`` / ``models.*`` routines the drop-in objects provide."""
    from wmf_b200 import CuModule, ModelsModule
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tree = ast.parse(open(WMF_PY, encoding="latin-1").read())
    objs = {"cu": CuModule, "models": ModelsModule}
    classes = {n.name: n for n in ast.walk(tree) if isinstance(n, ast.ClassDef)}
    methods = {}
    for cname in ("Basin", "SimuBasin"):                      # SimuBasin overrides Basin
        for fn in classes[cname].body:
            if isinstance(fn, ast.FunctionDef):
                methods[fn.name] = fn
    roots = ["__init__", "set_Geomorphology", "set_Control", "run_shia", "rain_interpolate_idw", "rain_interpolate_mit",
             "set_Storage", "set_PhysicVariables", "set_Speed_type", "set_Sediments", "set_Slides", "set_Floods"]
    seen, stack = set(), [r for r in roots if r in methods]
    while stack:
        m = stack.pop()
        if m in seen or m not in methods:
            continue
        seen.add(m)
        for c in ast.walk(methods[m]):
            if (isinstance(c, ast.Call) and isinstance(c.func, ast.Attribute) and isinstance(c.func.value, ast.Name)
                    and c.func.value.id == "self"):
                stack.append(c.func.attr)
    assert {"Points_Points2Stream", "GetGeo_Cell_Basics", "Transform_Map2Basin"} <= seen
    missing = []
    for m in sorted(seen):
        for c in ast.walk(methods[m]):
            if (isinstance(c, ast.Call) and isinstance(c.func, ast.Attribute) and isinstance(c.func.value, ast.Name)
                    and c.func.value.id in objs and getattr(objs[c.func.value.id], c.func.attr, None) is None):
                missing.append(f"wmf.py:{c.lineno} {m}: {c.func.value.id}.{c.func.attr}")
    assert not missing, "\n".join(missing)
