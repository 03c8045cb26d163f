"""Repository invariants the tier contract states: the product never touches the oracle, never reads the reference at
run time, has no CPU fallback switch, and avoids the CUDA batch-memcpy entry points this pool forbids."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deeplearningrecommendationsystem_b200")


def _files(top, exts):
    for d, _, names in os.walk(top):
        if "build" in d.split(os.sep) or "__pycache__" in d:
            continue
        for n in names:
            if n.endswith(exts):
                yield os.path.join(d, n)


def test_product_never_imports_the_oracle():
    bad = []
    for p in _files(PKG, (".py",)):
        src = open(p).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
            bad.append(p)
    assert not bad, f"product files import oracle/: {bad}"


def test_only_tests_smoke_and_bench_use_the_oracle():
    users = []
    for p in _files(ROOT, (".py",)):
        rel = os.path.relpath(p, ROOT)
        if rel.startswith(("oracle" + os.sep, "tests" + os.sep)):
            continue
        if re.search(r"^\s*(from|import)\s+oracle\b", open(p).read(), flags=re.M):
            users.append(rel)
    assert sorted(users) == ["__graft_entry__.py", "bench.py"], users


def test_nothing_reads_the_reference_tree_at_run_time():
    for p in list(_files(PKG, (".py", ".cu", ".cuh"))) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        assert "/root/reference" not in open(p).read(), p


def test_forbidden_batch_memcpy_entry_points_are_absent():
    names = ["cudaMemcpy" + "BatchAsync", "cudaMemcpy3D" + "BatchAsync", "cuMemcpy" + "BatchAsync", "cuMemcpy3D" + "BatchAsync"]
    for p in _files(ROOT, (".py", ".cu", ".cuh", ".h")):
        src = open(p).read()
        for n in names:
            assert n not in src, (p, n)


def test_every_kernel_source_targets_sm100a_only():
    build = open(os.path.join(PKG, "build.py")).read()
    assert "arch=compute_100a,code=sm_100a" in build and "-lineinfo" in build
    assert not re.search(r"sm_(7|8|9)\d", build)
