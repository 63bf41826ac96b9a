"""CPU-only checks of the drop-in boundary: the C-ABI library loads without a GPU, exports exactly what
include/plssvm_b200.h declares, fails loudly without a device, and the product never touches the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

import plssvm_b200 as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "plssvm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plssvm_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = pb.load_library()
    declared = _header_symbols()
    assert sorted(pb.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/plssvm_b200.h but not exported"


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "plssvm_b200.h")).read()
    for cite in ("csvm.hpp:188-208", "gpu_csvm.hpp:477-654", "gpu_csvm.hpp:656-730", "csvm.cu:110-129", "csvm.cu:134-153", "csvm.cu:158-165", "csvm.cu:170-186"):
        assert cite in text


def test_header_documents_every_option_the_library_accepts():
    """Every key plssvm_b200_set_option accepts (backend.cu) is described in the header's comment on that entry point, and so is every value of "impl"
    the default build accepts."""
    src = open(os.path.join(ROOT, "plssvm_b200", "csrc", "backend.cu")).read()
    header = open(os.path.join(ROOT, "include", "plssvm_b200.h")).read()
    keys = sorted(set(re.findall(r'\bk [!=]= "([a-z_0-9]+)"', src)))
    assert len(keys) >= 15 and "impl" in keys and "fp32_pair" in keys
    for k in keys:
        assert f'"{k}"' in header, f'option "{k}" is accepted by plssvm_b200_set_option but not documented in include/plssvm_b200.h'
    known = re.search(r"const bool known = ([^;]+);", src).group(1).split("(EXPERIMENTAL")[0]
    for v in re.findall(r"value == (\d+)", known):
        assert re.search(rf"\b{v}\b", header[header.index("tuning / debugging knobs"):]), f"impl = {v} not documented"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pb.BackendError) as e:
        pb.Backend(0)
    assert e.value.code == 2  # PLSSVM_B200_ERR_CUDA


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "plssvm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "lssvm_oracle" not in src, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read()


@pytest.mark.parametrize("T", [1, 2, 5, 11, 12, 13, 24, 25, 40, 100, 513])
def test_triangle_order_is_a_bijection(T):
    total = pb.tri_num_tiles(T)
    assert total == T * (T + 1) // 2
    seen = set()
    for L in range(total):
        I, J = pb.tri_decode(T, L)
        assert 0 <= J <= I < T
        assert pb.tri_encode(T, I, J) == L
        seen.add((I, J))
    assert len(seen) == total


def test_triangle_order_is_banded():
    """Consecutive tiles stay inside one band of 12 tile rows (L2 reuse), bands are visited top to bottom."""
    T = 100
    prev_band = 0
    for L in range(pb.tri_num_tiles(T)):
        I, _ = pb.tri_decode(T, L)
        assert I // 12 >= prev_band
        prev_band = I // 12


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_rank_ranges_partition_the_triangle_evenly(world):
    total = pb.tri_num_tiles(512)
    ranges = [pb.rank_range(total, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == total
    for (a, b), (c, d) in zip(ranges, ranges[1:]):
        assert b == c
    sizes = [hi - lo for lo, hi in ranges]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("planes,box_rows,slab", [(7, 128, 64), (7, 64, 64), (3, 128, 64), (3, 64, 64), (3, 128, 32), (4, 128, 32)])
def test_int8_plane_layout_is_a_bijection_of_swizzled_boxes(planes, box_rows, slab):
    """The digit planes are stored as contiguous operand boxes that already are the shared-memory image tcgen05.mma expects:
    box (R, ks) = planes x box_rows x slab bytes, plane-major; inside a plane K-major rows of `slab` bytes whose 16-byte chunks are XOR-swizzled
    with the address bits of the hardware mode — SWIZZLE_64B: bits 1..2 of the row, SWIZZLE_32B (fp32 kernel): bit 2 of the row.
    Host-only: the offset function is the single source of truth of kernel and split."""
    slabs, rows = 3, 2 * box_rows
    seen = set()
    for r in range(rows):
        for k in range(0, slab * slabs, 4):  # the split kernel writes 4 consecutive features per store
            for p in range(planes):
                off = pb.i8_plane_offset(r, k, p, planes, box_rows, slabs, slab)
                assert off % 4 == 0 and off not in seen
                seen.add(off)
                box, within = divmod(off, planes * box_rows * slab)
                assert box == (r // box_rows) * slabs + k // slab
                plane, in_plane = divmod(within, box_rows * slab)
                assert plane == p
                rr, byte = divmod(in_plane, slab)
                assert rr == r % box_rows
                sw = (rr // 2) % 4 if slab == 64 else (rr // 4) % 2
                assert byte == ((((k % slab) // 16) ^ sw) * 16 + k % 16)
                # the swizzle is the hardware's: XOR of address bits [4, 4 + log2(slab / 16)) with bits [7, 7 + log2(slab / 16)) of the in-plane address
                linear = rr * slab + k % slab
                mask = (slab // 16 - 1) << 4
                assert in_plane == linear ^ (((linear >> 7) << 4) & mask)
    assert len(seen) == rows * (slab // 4) * slabs * planes and max(seen) < planes * rows * slab * slabs


def test_tile_size_matches_design():
    assert pb.tile_size() == 128


def test_kernel_ids_follow_the_reference_enum():
    # include/plssvm/kernel_function_types.hpp:31-38
    assert (pb.kernel_id("linear"), pb.kernel_id("polynomial"), pb.kernel_id("rbf")) == (0, 1, 2)
    with pytest.raises(ValueError):
        pb.kernel_id("sigmoid")
