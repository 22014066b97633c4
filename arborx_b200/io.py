"""Point-cloud files of the reference's benchmarks (benchmarks/cluster/data_timpl.hpp:65-143): `[num_points][dim]`
followed by num_points x dim float32 values, as text or as raw binary (int32 header).  Distributed reading splits
the points evenly over the ranks, the remainder going to the last one, exactly like the reference."""
import numpy as np


def load_points(filename, binary=True, max_num_points=-1, comm_rank=0, comm_size=1, dim=3):
    """-> float32 [n, dim] (this rank's share)."""
    if comm_size > 1 and not binary:
        raise RuntimeError("Distributed reading only works with binary files")
    if binary:
        with open(filename, "rb") as f:
            hdr = np.fromfile(f, np.int32, 2)
            if hdr.size != 2:
                raise ValueError("%s: truncated header" % filename)
            num_points, file_dim = int(hdr[0]), int(hdr[1])
            if file_dim != dim:
                raise ValueError("%s holds %d-D points, %d-D expected" % (filename, file_dim, dim))
            if 0 < max_num_points < num_points:
                num_points = max_num_points
            per = num_points // comm_size
            mine = per + ((num_points % per) if (comm_rank == comm_size - 1 and per) else 0)
            f.seek(per * comm_rank * dim * 4, 1)
            data = np.fromfile(f, np.float32, mine * dim)
        if data.size != mine * dim:
            raise ValueError("%s: truncated data" % filename)
        return data.reshape(mine, dim)
    tokens = open(filename).read().split()
    num_points, file_dim = int(tokens[0]), int(tokens[1])
    if file_dim != dim:
        raise ValueError("%s holds %d-D points, %d-D expected" % (filename, file_dim, dim))
    if 0 < max_num_points < num_points:
        num_points = max_num_points
    vals = np.array(tokens[2:2 + num_points * dim], np.float32)
    if vals.size != num_points * dim:
        raise ValueError("%s: truncated data" % filename)
    return vals.reshape(num_points, dim)


def save_points(filename, points, binary=True):
    p = np.ascontiguousarray(points, np.float32)
    if binary:
        with open(filename, "wb") as f:
            np.array([p.shape[0], p.shape[1]], np.int32).tofile(f)
            p.tofile(f)
    else:
        with open(filename, "w") as f:
            f.write("%d %d\n" % p.shape)
            for row in p:
                f.write(" ".join(repr(float(x)) for x in row) + "\n")
