"""The reference's point-file formats (benchmarks/cluster/data_timpl.hpp:65-143): text / binary round trips, the
distributed split, and the benchmark's own 8-point input (benchmarks/cluster/input.txt restated)."""
import numpy as np
import pytest

from arborx_b200 import io
from tests import clouds

INPUT_TXT = "8 3\n0 0 0\n1 1 1\n2 2 2\n3 3 3\n9 9 9\n10 10 10\n11 11 11\n20 20 20\n"


def test_text_and_binary_round_trip(tmp_path):
    pts = clouds.filled_box(5, 1001)
    for binary in (True, False):
        fn = str(tmp_path / ("p.bin" if binary else "p.txt"))
        io.save_points(fn, pts, binary)
        assert np.array_equal(io.load_points(fn, binary), pts)
        assert np.array_equal(io.load_points(fn, binary, max_num_points=10), pts[:10])
    txt = tmp_path / "input.txt"
    txt.write_text(INPUT_TXT)
    p = io.load_points(str(txt), binary=False)
    assert p.shape == (8, 3) and p[4, 0] == 9


def test_distributed_split(tmp_path):
    pts = clouds.filled_box(6, 1003)
    fn = str(tmp_path / "p.bin")
    io.save_points(fn, pts)
    parts = [io.load_points(fn, True, comm_rank=r, comm_size=4) for r in range(4)]
    assert [len(p) for p in parts] == [250, 250, 250, 253]
    assert np.array_equal(np.concatenate(parts), pts)
    with pytest.raises(RuntimeError):
        io.load_points(fn, False, comm_rank=0, comm_size=2)
    with pytest.raises(ValueError):
        io.load_points(fn, True, dim=2)
