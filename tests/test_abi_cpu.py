"""CPU tests of the C-ABI boundary (no compute): the library loads, exports every symbol that
include/lichee_b200.h declares, mirrors the struct layouts, validates its arguments and refuses to
run without a GPU (there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "lichee_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lichee_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    import lichee_b200
    lib = lichee_b200.lib()
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(lichee_b200.EXPORTED_SYMBOLS) == names
    assert b"sm_100a" in lib.lichee_version()


def test_struct_layouts_match_the_header(built):
    import lichee_b200
    assert ctypes.sizeof(lichee_b200.SearchParams) == 8 + 4 + 4 + 8 + 8 + 4 + 4 + 8 + 32
    assert ctypes.sizeof(lichee_b200.SearchStats) == 6 * 8 + 8 + 3 * 5 * 8 + 8 + 3 * 8
    p = lichee_b200.default_params()
    assert (p.vaf_error_margin, p.num_save, p.max_num_trees, p.max_num_grow_calls) == (0.1, 1, 100000, 100000000)


def test_argument_validation_needs_no_gpu(built):
    import lichee_b200
    with pytest.raises(lichee_b200.LicheeError) as e:            # NaN centroid
        lichee_b200.LineageSearch([[1], []], np.array([[0.5], [float("nan")]]))
    assert e.value.code == -1
    with pytest.raises(lichee_b200.LicheeError) as e:            # adjacency lists / matrix disagree
        lichee_b200.LineageSearch([[1], [], []], np.array([[0.5], [0.2]]))
    assert e.value.code == -1
    with pytest.raises(lichee_b200.LicheeError) as e:            # too many samples
        lichee_b200.LineageSearch([[1], []], np.zeros((2, 65)))
    assert e.value.code == -4
    with pytest.raises(lichee_b200.LicheeError) as e:            # too many nodes
        lichee_b200.LineageSearch([[] for _ in range(129)], np.zeros((129, 2)))
    assert e.value.code == -4
    with pytest.raises(lichee_b200.LicheeError) as e:            # -s out of range
        lichee_b200.LineageSearch([[1], []], np.array([[0.5], [0.2]]), num_save=5000)
    assert e.value.code == -1
    with pytest.raises(lichee_b200.LicheeError) as e:            # edge target out of range
        lichee_b200.LineageSearch([[7], []], np.array([[0.5], [0.2]]))
    assert e.value.code == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(built):
    import lichee_b200
    assert lichee_b200.lib().lichee_device_count() == 0
    with pytest.raises(lichee_b200.LicheeError) as e:
        lichee_b200.LineageSearch([[1], []], np.array([[0.5], [0.2]]))
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "lichee_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "lichee_oracle" not in txt, f
