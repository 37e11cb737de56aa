"""The product never routes through the oracle or a CPU path: static check of the imports, and every operator wrapper
refuses CPU tensors (the -m gpu suite checks the same with the library loaded)."""
import ast
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _imports(path):
    tree = ast.parse(open(path).read(), path)
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name
        elif isinstance(node, ast.ImportFrom) and node.module:
            yield node.module if node.level == 0 else "." * node.level + node.module


def test_only_tests_smoke_and_bench_cpu_legs_import_the_oracle():
    offenders = []
    for base in ("univs_b200", "tools"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(".py"):
                    p = os.path.join(d, f)
                    if any(m == "oracle" or m.startswith("oracle.") for m in _imports(p)):
                        offenders.append(os.path.relpath(p, ROOT))
    assert offenders == []
    # bench.py: the oracle is imported inside run_cpu_reference only (cpu_baseline / --impl reference legs);
    # __graft_entry__.py: inside smoke() only
    for fname, func in (("bench.py", "run_cpu_reference"), ("__graft_entry__.py", "smoke")):
        tree = ast.parse(open(os.path.join(ROOT, fname)).read())
        for node in tree.body:
            inner = list(ast.walk(node))
            uses = [n for n in inner if isinstance(n, (ast.Import, ast.ImportFrom))
                    and any("oracle" in (getattr(n, "module", None) or "") or "oracle" in a.name for a in n.names)]
            if uses:
                assert isinstance(node, ast.FunctionDef) and node.name == func, (fname, getattr(node, "name", node))


def test_operator_wrappers_refuse_cpu_tensors():
    from univs_b200 import ops
    from univs_b200._cabi import UnivsB200Error
    x = torch.zeros(1, 4, 256)
    with pytest.raises(UnivsB200Error):
        ops.mask_einsum(x, torch.zeros(1, 16, 256), mode="tf32")
    with pytest.raises(UnivsB200Error):
        ops.mha_core(x, x, x)
    with pytest.raises(UnivsB200Error):
        ops.swin_window_attention(torch.zeros(1, 7, 7, 96), torch.zeros(96), torch.zeros(169, 1), 1, 7, 0)
    with pytest.raises(UnivsB200Error):
        ops.ms_deform_attn_encoder(torch.zeros(1, 4, 8, 32), [(2, 2)], [0], torch.zeros(1, 4, 8 * 1 * 4 * 3), 1, 4)
