"""CPU tier: the closed form behind the lattice symbolic pass (csrc/femx_pattern.cu: lat_locate / lat_prefix, through the
host entry point femx_lattice_prefix).  On the device it replaces the row-length array, the scan over the rows and the
compaction of the rows outside the class; here it is checked against a plain running sum over every node id, on lattices
whose node numbering has holes in front, between lines, between planes and behind."""
import numpy as np
import pytest

import femx


def brute(cells, strides, node0, w, n_nodes):
    dim = len(cells)
    cn = list(cells) + [1] * (3 - dim)
    s = list(strides) + [0] * (3 - dim)
    out = np.zeros(n_nodes + 1, np.int64)
    acc = 0
    for node in range(n_nodes + 1):
        out[node] = acc
        q = node - node0
        if q < 0:
            continue
        k = q // s[2] if dim == 3 else 0
        q -= k * s[2] if dim == 3 else 0
        j, i = q // s[1], q % s[1]
        if i > cn[0] or j > cn[1] or k > cn[2] or (dim == 2 and k):
            continue
        cls = lambda p, c: 0 if p == 0 else (2 if p == c else 1)
        acc += w[cls(i, cn[0]) + 3 * cls(j, cn[1]) + (9 * cls(k, cn[2]) if dim == 3 else 0)]
    return out


@pytest.mark.parametrize("seed", range(12))
def test_closed_form_prefix_equals_running_sum(seed):
    rng = np.random.default_rng(seed)
    dim = 2 + seed % 2
    cells = [int(rng.integers(1, 6)) for _ in range(dim)]
    sy = cells[0] + 1 + int(rng.integers(0, 3))
    strides = [1, sy] + ([(cells[1] + 1) * sy + int(rng.integers(0, 4))] if dim == 3 else [])
    node0 = int(rng.integers(0, 5))
    w = rng.integers(0, 30, 27)
    if dim == 2:
        w[9:] = 0
    last = node0 + cells[0] + cells[1] * strides[1] + (cells[2] * strides[2] if dim == 3 else 0)
    n_nodes = last + 1 + int(rng.integers(0, 6))
    ref = brute(cells, strides, node0, w, n_nodes)
    got = femx.lattice_prefix(cells, strides, node0, w, np.arange(n_nodes + 1))
    assert np.array_equal(got, ref)


def test_row_pointer_of_a_kuhn_box():
    """weights = row length per class: the prefix is the CSR row pointer of femx_mesh_box's pattern (oracle)."""
    from oracle import oracle as orc
    nx, ny, nz = 4, 3, 5
    X, Y, Z, conn = orc.box_mesh(nx, ny, nz)
    rp, ci = orc.pattern(conn, len(X))
    # row length per class from one representative node of each class
    w = np.zeros(27, np.int64)
    for node in range(len(X)):
        i, j, k = node % (nx + 1), (node // (nx + 1)) % (ny + 1), node // ((nx + 1) * (ny + 1))
        c = lambda p, n: 0 if p == 0 else (2 if p == n else 1)
        w[c(i, nx) + 3 * c(j, ny) + 9 * c(k, nz)] = rp[node + 1] - rp[node]
    got = femx.lattice_prefix([nx, ny, nz], [1, nx + 1, (nx + 1) * (ny + 1)], 0, w, np.arange(len(X) + 1))
    assert np.array_equal(got, rp)


def test_bad_arguments_are_rejected():
    with pytest.raises(femx.FemxError):
        femx.lattice_prefix([4, 3], [1, 3], 0, np.zeros(27, int), [0])     # stride narrower than the line
