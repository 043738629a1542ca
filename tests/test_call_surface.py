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
