"""The C-ABI library loads (no GPU needed) and exports exactly what include/ditb200.h declares."""
import ctypes
import os
import re

from util import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "ditb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ditb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from fast_dit_b200 import _lib

    names = _declared()
    assert len(names) >= 18
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ditb200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "python binding and header disagree"
    assert lib.ditb200_abi_version() == _lib.ABI_VERSION == 3


def test_struct_layouts_match_header():
    from fast_dit_b200 import _lib

    # field counts of the three argument structs as written in the header
    src = open(os.path.join(ROOT, "include", "ditb200.h")).read()
    for cname, cls in [("ditb200_gemm_args", _lib.GemmArgs), ("ditb200_step_args", _lib.StepArgs),
                       ("ditb200_loss_args", _lib.LossArgs)]:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = names[0].split()[-1].lstrip("*")
            fields.append(first)
            fields += [n.strip().lstrip("*") for n in names[1:]]
        assert fields == [f[0] for f in cls._fields_], cname


def test_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    from fast_dit_b200 import _lib

    lib = _lib.load()
    rc = lib.ditb200_ln_modulate(None, None, None, 0, None, 0, None, 1, 1, 4, ctypes.c_float(1e-6), 0, None)
    assert rc == -1
    assert b"null pointer" in lib.ditb200_last_error()
    args = _lib.GemmArgs()
    assert lib.ditb200_gemm(ctypes.byref(args), None) == -1


def test_no_fallback_on_cpu_tensors():
    import pytest
    import torch

    from fast_dit_b200 import _lib, ops

    with pytest.raises(_lib.Ditb200Error):
        ops.timestep_embedding(torch.zeros(2, dtype=torch.long), 256)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fast_dit_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)


def test_ptxas_spill_budget_of_the_hot_kernels():
    """Build hygiene, read from the `-Xptxas -v` logs the in-tree build leaves next to the objects: every block of
    these kernels has 10 warps, which caps them at 168 registers per thread, and what does not fit goes to local
    memory.  The single-tile attention forward must not spill at all (its softmax loop is the kernel); the others
    are pinned at what the measured builds had, so a change that pushes an elementwise warp into local memory — the
    backward kernel went from 24 to 152 bytes and from 101 to 115 us when its issue loops were rewritten — fails here
    before it costs a GPU call.  The tcgen05 GEMM's ~500 bytes sit in its register-path epilogues."""
    import glob

    import pytest

    logs = glob.glob(os.path.join(ROOT, "fast_dit_b200", "lib", "obj", "*.log"))
    if not logs:
        pytest.skip("no ptxas logs: the library was not built in this tree")
    spills = {}
    for path in logs:
        name = None
        for line in open(path):
            m = re.search(r"Function properties for (\S+)", line)
            if m:
                name = m.group(1)
            m = re.search(r"(\d+) bytes spill stores", line)
            if m and name:
                spills[name] = int(m.group(1))
    budget = {"18attn_fwd_tc_kernelE": 0, "21attn_fwd_tc_kv_kernelE": 64, "18attn_bwd_tc_kernelE": 24,
              "14gemm_tc_kernelI": 520, "18ln_modulate_kernel": 0, "24ln_modulate_resid_kernel": 0,
              "20p_sample_step_kernel": 0}
    for key, limit in budget.items():
        hit = {k: v for k, v in spills.items() if key in k}
        assert hit, key
        for k, v in hit.items():
            assert v <= limit, (k, v, limit)
