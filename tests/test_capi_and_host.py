"""CPU: the C-ABI library builds, loads and exports every symbol include/nalgebra_b200.h declares;
the host-side mirror behaves like the reference on shapes/errors; no compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nalgebra_b200.h")).read()
    return sorted(set(re.findall(r"NAB_API\s+[\w\s\*]+?\b(na_\w+|d[a-z0-9]+_)\s*\(", src)))


def test_header_symbols_are_exported(nab):
    from nalgebra_b200 import _capi
    lib = _capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 60 and "dgetrf_" in declared and "na_dgemm" in declared
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"not exported: {missing}"
    # the ctypes signature table covers exactly the header
    assert sorted(_capi.SIGNATURES) == declared
    assert lib.na_version().decode().startswith("nalgebra_b200")


def test_library_has_no_torch_or_oracle_dependency(nab):
    from nalgebra_b200 import _capi
    import subprocess
    out = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "libcudart" not in out   # static cudart, plain C ABI


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nalgebra_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "nalgebra_oracle" not in txt, f


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="needs a machine WITHOUT a GPU")
def test_compute_fails_loudly_without_gpu(nab):
    from nalgebra_b200 import _capi
    a = np.ones((8, 8), order="F")
    with pytest.raises(_capi.NalgebraB200Error) as e:
        nab.gemm(1.0, a, a, 0.0, a.copy(order="F"))
    assert e.value.status == _capi.NA_ECUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(_capi.NalgebraB200Error):
        nab.Cholesky.new(np.eye(4))
    with pytest.raises(_capi.NalgebraB200Error):
        nab.LU.new(np.eye(4))
    with pytest.raises(_capi.NalgebraB200Error):
        nab.QR.new(np.eye(4))


def test_shape_errors_mirror_reference_panics(nab):
    a = np.ones((3, 4), order="F")
    with pytest.raises(ValueError, match="multiplication"):
        nab.gemm(1.0, a, a, 0.0, np.zeros((3, 4), order="F"))
    with pytest.raises(ValueError, match="addition"):
        nab.gemm(1.0, a, a.T, 0.0, np.zeros((4, 4), order="F"))
    with pytest.raises(ValueError, match="square"):
        nab.Cholesky.new(a)
    with pytest.raises(ValueError, match="non-square"):
        nab.LU(a, nab.PermutationSequence.identity(3)).solve(np.ones((3, 1)))


def test_permutation_sequence_semantics(nab, oracle):
    p = nab.PermutationSequence.identity(4)
    p.append_permutation(0, 2); p.append_permutation(1, 1); p.append_permutation(2, 3)
    assert len(p) == 2 and p.determinant() == 1.0
    m = np.arange(16.0).reshape(4, 4)
    x = m.copy(); p.permute_rows(x)
    assert np.array_equal(x, oracle.permute_rows(p.ipiv, m))
    p.inv_permute_rows(x)
    assert np.array_equal(x, m)
    p.append_permutation(0, 1)
    assert p.determinant() == -1.0
    with pytest.raises(ValueError):
        q = nab.PermutationSequence.identity(1); q.append_permutation(0, 1); q.append_permutation(0, 1)


def test_sharding_grid_arithmetic():
    from nalgebra_b200 import sharding as S
    assert [S.process_grid(w) for w in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    for world in (1, 2, 3, 4, 8):
        cover = np.zeros((37, 53), dtype=int)
        for r in range(world):
            r0, r1, c0, c1 = S.gemm_tile(r, world, 37, 53)
            cover[r0:r1, c0:c1] += 1
        assert (cover == 1).all()
    assert S.block_cyclic_local_blocks(1, 4, 10) == [1, 5, 9] and S.block_cyclic_owner(7, 4) == 3


def test_gemm_exchange_piece_order():
    """K-piece schedule of the multi-GPU GEMM step (bench.py): every piece exactly once, local pieces before the ones
    that wait for a gather, and the classes agree with the ownership rule of the non-replicated layout."""
    from nalgebra_b200 import sharding as S
    n = 16384
    for world in (2, 4, 8):
        pr, pc = S.process_grid(world)
        kca, kcb = n // pc, n // pr
        kp = min(kca, kcb, 2048)
        n_local_local = 0
        for rank in range(world):
            r, c = divmod(rank, pc)
            order = S.gemm_piece_order(n, kp, kca, kcb, r, c)
            assert sorted(order) == list(range(n // kp))
            classes = [S.gemm_piece_class(t, kp, kca, kcb, r, c) for t in order]
            assert classes == sorted(classes)
            for t in order:
                k0 = t * kp
                a_local = c * kca <= k0 < (c + 1) * kca
                b_local = r * kcb <= k0 < (r + 1) * kcb
                assert S.gemm_piece_class(t, kp, kca, kcb, r, c) == (0 if a_local and b_local else 1 if a_local else 2 if b_local else 3)
            # every rank owns one A chunk and one B chunk: kca / kp pieces with A local, kcb / kp with B local
            assert sum(1 for x in classes if x in (0, 1)) == kca // kp
            assert sum(1 for x in classes if x in (0, 2)) == kcb // kp
            n_local_local += classes.count(0)
        assert n_local_local > 0



def test_rust_sys_crate_matches_header():
    """rust/nalgebra-b200-sys/src/lib.rs cannot be compiled here (no cargo/rustc): at least every function it declares
    must exist in include/nalgebra_b200.h with the same number of arguments."""
    hdr = open(os.path.join(ROOT, "include", "nalgebra_b200.h")).read()
    rs = open(os.path.join(ROOT, "rust", "nalgebra-b200-sys", "src", "lib.rs")).read()
    decl = {m.group(1): m.group(2) for m in re.finditer(r"NAB_API\s+[\w\s\*]+?\b(na_\w+)\s*\(([^;]*?)\)\s*;", hdr, re.S)}
    fns = re.findall(r"pub fn (na_\w+)\s*\(([^;]*?)\)\s*(?:->\s*[\w\*\s]+)?;", rs, re.S)
    assert len(fns) >= 20
    for name, args in fns:
        assert name in decl, name
        n_rs = 0 if not args.strip() else args.count(":")
        c_args = decl[name].strip()
        n_c = 0 if c_args in ("", "void") else c_args.count(",") + 1
        assert n_rs == n_c, (name, n_rs, n_c)
