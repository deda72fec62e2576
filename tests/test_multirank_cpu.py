"""World-size-2 (and 4) gloo tests of the multi-GPU host logic on CPU: the
Z-curve partition, interior/boundary ordering, halo slot numbering and the
point-to-point exchange schedule (spectre_b200.domain.Partition and
spectre_b200.evolution.HaloExchange).  The face packing that the pack_halo
kernel does on the GPU is emulated with numpy here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spectre_b200 import domain
from spectre_b200.evolution import HaloExchange

N = 3
F = N * N
HC = 7  # components per face point in this test


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _face_points(d):
    dim, side = d // 2, d % 2
    fixed = N - 1 if side else 0
    q = np.arange(F)
    a, b = q % N, q // N
    if dim == 0:
        return fixed + N * (a + N * b)
    if dim == 1:
        return a + N * (fixed + N * b)
    return a + N * (b + N * fixed)


def _field(global_element, n_elements):
    """Deterministic per-element data [HC, n] that identifies element and point."""
    n = N ** 3
    c = np.arange(HC)[:, None]
    p = np.arange(n)[None, :]
    return global_element * 1000.0 + c * 100.0 + p


def _worker(rank, world, port, refinement, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    brick = domain.Brick([0, 0, 0], [1, 1, 1], refinement, N)
    nb = brick.neighbors()
    part = domain.Partition(nb, world, rank)
    # local ordering: interior first, boundary last; ghosts only on boundary elements
    ln = part.local_neighbors
    assert (ln[:part.n_interior] >= -1).all()
    assert ((ln[part.n_interior:] <= -2).any(axis=1)).all()
    # pack (numpy emulation of pack_halo_kernel)
    send = torch.zeros(part.n_ghost * HC * F, dtype=torch.float64)
    sv = send.numpy().reshape(part.n_ghost, HC, F)
    for slot, (le, d) in enumerate(part.send_map):
        sv[slot] = _field(part.global_ids[le], brick.n_elements)[:, _face_points(d)]
    recv = torch.full((part.n_ghost * HC * F,), -1.0, dtype=torch.float64)
    halo = HaloExchange(part, HC * F, dist)
    for w in halo.start(send, recv):
        w.wait()
    rv = recv.numpy().reshape(part.n_ghost, HC, F)
    # every ghost slot must hold the face of the true neighbour, seen from its side
    checked = 0
    for le in range(part.n_local):
        for d in range(6):
            v = ln[le, d]
            if v <= -2:
                g_nb = nb[part.global_ids[le], d]
                expect = _field(g_nb, brick.n_elements)[:, _face_points(d ^ 1)]
                np.testing.assert_array_equal(rv[-(v + 2)], expect)
                checked += 1
            elif v >= 0:
                assert part.global_ids[v] == nb[part.global_ids[le], d]
    assert checked == part.n_ghost
    # partitions tile the element list
    counts = [None] * world
    dist.all_gather_object(counts, (part.n_local, sorted(part.global_ids.tolist())))
    if rank == 0:
        all_ids = sorted(sum((c[1] for c in counts), []))
        assert all_ids == list(range(brick.n_elements))
        assert max(c[0] for c in counts) - min(c[0] for c in counts) <= 1
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,refinement", [(2, [1, 1, 1]), (2, [2, 1, 1]), (4, [2, 2, 1])])
def test_partition_and_halo_exchange_gloo(world, refinement):
    mp.spawn(_worker, args=(world, _free_port(), refinement, None), nprocs=world, join=True)


def test_partition_single_rank_has_no_ghosts():
    brick = domain.Brick([0, 0, 0], [1, 1, 1], [2, 2, 2], 4)
    part = domain.Partition(brick.neighbors(), 1, 0)
    assert part.n_ghost == 0 and part.n_interior == part.n_local == 64
    np.testing.assert_array_equal(part.local_neighbors, brick.neighbors())


def test_z_curve_order_matches_reference_bit_interleave():
    """ZCurve.cpp:17-80: for equal refinement the index is the Morton code
    with x as the least significant bit."""
    for (ix, iy, iz) in [(1, 0, 0), (0, 1, 0), (0, 0, 1), (3, 2, 1), (5, 7, 2)]:
        want = 0
        for b in range(3):
            want |= ((ix >> b) & 1) << (3 * b) | ((iy >> b) & 1) << (3 * b + 1) | \
                ((iz >> b) & 1) << (3 * b + 2)
        assert domain.z_curve_index(ix, iy, iz, (3, 3, 3)) == want
    # unequal refinement: dimensions with level 0 are skipped
    assert domain.z_curve_index(1, 0, 1, (1, 0, 1)) == 0b11
    assert domain.z_curve_index(2, 0, 1, (2, 0, 1)) == 0b101
    for (ix, iz), want in ZCURVE_KNOWN_ANSWERS.items():
        assert domain.z_curve_index(ix, 0, iz, (2, 0, 3)) == want


# tests/Unit/Domain/Structure/Test_ZCurve.cpp:236-291: refinement levels (2, 0, 3),
# (x index, z index) -> Z-curve index
ZCURVE_KNOWN_ANSWERS = {
    (0, 0): 0, (1, 0): 1, (2, 0): 4, (3, 0): 5, (0, 1): 2, (1, 1): 3, (2, 1): 6, (3, 1): 7,
    (0, 2): 8, (1, 2): 9, (2, 2): 12, (3, 2): 13, (0, 3): 10, (1, 3): 11, (2, 3): 14, (3, 3): 15,
    (0, 4): 16, (1, 4): 17, (2, 4): 20, (3, 4): 21, (0, 5): 18, (1, 5): 19, (2, 5): 22,
    (3, 5): 23, (0, 6): 24}


def test_cpp_z_curve_index_shim():
    """domain::z_curve_index of SpectreShims.hpp against the reference's known answers and
    against spectre_b200.domain.z_curve_index on random unequal refinements."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "_build", "orientation_codes")
    src = os.path.join(root, "tests", "helpers", "orientation_codes.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-o", exe, src, "-L",
                           os.path.join(root, "spectre_b200"), "-ldgrhs",
                           "-Wl,-rpath," + os.path.join(root, "spectre_b200")])
    lines, want = [], []
    for (ix, iz), v in ZCURVE_KNOWN_ANSWERS.items():
        lines.append(f"Z 2 {ix} 0 0 3 {iz}")
        want.append(v)
    rng = np.random.default_rng(3)
    for _ in range(200):
        lev = rng.integers(0, 5, 3)
        idx = [int(rng.integers(0, 2 ** l)) for l in lev]
        lines.append("Z " + " ".join(f"{l} {i}" for l, i in zip(lev, idx)))
        want.append(domain.z_curve_index(*idx, tuple(int(l) for l in lev)))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert [int(v) for v in out.stdout.split()] == want


def _shell_worker(rank, world, port, order, results):
    """The same exchange on the six-wedge shell: neighbours across wedge
    boundaries are not aligned, so a ghost slot holds the SENDER's face in the
    sender's ordering and the receiver's orientation table (neighbour direction
    + face permutation) finds the matching point -- what the face kernel does."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if order == "bco":
        # BASELINE configs[4]: the 44-block BinaryCompactObject domain cut in block order
        from spectre_b200 import bco
        sh = bco.BinaryCompactObject(8.0, -8.0, 0.8, 4.0, 0.8, 4.0, 60.0, 300.0, 0, N,
                                     opening_angle_degrees=120.0)
    else:
        sh = domain.SphericalShell(1.9, 4.0, (1, 1), N, order=order)
    nb, (nd, perm) = sh.neighbors(), sh.neighbor_orientations()
    x = sh.coords()
    part = domain.Partition(nb, world, rank, boundary_slots=True, neighbor_direction=nd,
                            face_permutation=perm)
    assert part.oriented
    ln = part.local_neighbors
    # "pack_halo": the face coordinates of the sender's face, in its own ordering
    send = torch.zeros(max(part.n_ghost, 1) * 3 * F, dtype=torch.float64)
    sv = send.numpy().reshape(-1, 3, F)
    for slot, (le, d) in enumerate(part.send_map):
        sv[slot] = x[part.global_ids[le]][:, _face_points(d)]
    recv = torch.full((max(part.n_ghost, 1) * 3 * F,), np.nan, dtype=torch.float64)
    halo = HaloExchange(part, 3 * F, dist)
    for w in halo.start(send, recv):
        w.wait()
    rv = recv.numpy().reshape(-1, 3, F)
    q = np.arange(F)
    qa, qb = q % N, q // N
    checked = 0
    for le in range(part.n_local):
        for d in range(6):
            v = ln[le, d]
            if v > -2 or -(v + 2) >= part.n_recv:
                continue  # local neighbour, or a boundary-condition slot
            code = part.local_face_permutation[le, d]
            na, nbb = np.where(code & 1, qb, qa), np.where(code & 1, qa, qb)
            if code & 2:
                na = N - 1 - na
            if code & 4:
                nbb = N - 1 - nbb
            mine = x[part.global_ids[le]][:, _face_points(d)]
            theirs = rv[-(v + 2)][:, na + N * nbb]
            np.testing.assert_allclose(theirs, mine, atol=1e-13 * max(1.0, np.abs(mine).max()))   # the same physical points
            checked += 1
    assert checked == part.n_recv
    counts = [None] * world
    dist.all_gather_object(counts, (part.n_local, part.n_recv))
    if rank == 0 and order == "radial":
        # cutting at constant radius: at most two spheres of faces per rank
        assert max(c[1] for c in counts) <= 2 * 6 * 4
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,order", [(2, "block"), (2, "radial"), (3, "block"), (2, "bco"),
                                         (4, "bco")])
def test_shell_partition_and_oriented_halo_exchange_gloo(world, order):
    mp.spawn(_shell_worker, args=(world, _free_port(), order, None), nprocs=world, join=True)


def _mortar_worker(rank, world, port, kind, results):
    """Mortars whose sides live on different ranks: the remote side's face arrives
    in the ghost slot named in Partition.local_mortars.  kind "shell": a shell with
    two wedges refined, i.e. mortar rows between blocks that are not aligned (fine
    direction | perm << 3) next to oriented conforming faces."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if kind == "brick":
        rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], N,
                                 {(0, 0, 0): (True, True, True), (1, 1, 0): (True, False, True)})
        nb, mt = rb.neighbors(), rb.mortars()
        part = domain.Partition(nb, world, rank, mortars=mt)
    else:
        rb = domain.SphericalShell(1.9, 2.9, [[(1, 1), (0, 0), (0, 0), (0, 0), (1, 0), (0, 0)]], N)
        nb, mt = rb.neighbors(), rb.mortars()
        assert ((mt[:, 3] >> 3) != 0).any()
        nd, perm = rb.neighbor_orientations()
        part = domain.Partition(nb, world, rank, neighbor_direction=nd, face_permutation=perm,
                                mortars=mt)
    send = torch.zeros(max(part.n_ghost, 1) * HC * F, dtype=torch.float64)
    sv = send.numpy().reshape(-1, HC, F)
    for slot, (le, d) in enumerate(part.send_map):
        sv[slot] = _field(part.global_ids[le], rb.n_elements)[:, _face_points(d)]
    recv = torch.full((max(part.n_ghost, 1) * HC * F,), -1.0, dtype=torch.float64)
    halo = HaloExchange(part, HC * F, dist)
    for w in halo.start(send, recv):
        w.wait()
    rv = recv.numpy().reshape(-1, HC, F)
    # global rows of my mortars, in the same order as local_mortars
    bounds = [(rb.n_elements * r) // world for r in range(world + 1)]
    owner = lambda g: max(r for r in range(world) if bounds[r] <= g)
    mine = [m for m in mt.tolist() if owner(m[0]) == rank or owner(m[2]) == rank]
    assert len(mine) == len(part.local_mortars)
    remote = 0
    for (ec, dc, ef, df, sa, sb), (lc, dc2, lf, df2, sa2, sb2) in zip(mine, part.local_mortars):
        assert (dc, df, sa, sb) == (dc2, df2, sa2, sb2)
        for g, d, l in ((ec, dc, lc), (ef, df & 7, lf)):
            if l >= 0:
                assert part.global_ids[l] == g
                assert part.local_neighbors[l, d] == domain.HANGING
                assert l >= part.n_interior or owner(ec) == owner(ef)
            else:
                expect = _field(g, rb.n_elements)[:, _face_points(d)]
                np.testing.assert_array_equal(rv[-(l + 2)], expect)
                remote += 1
    counts = [None] * world
    dist.all_gather_object(counts, remote)
    if rank == 0:
        assert sum(counts) > 0 and sum(counts) % 2 == 0   # every cut mortar is seen from both sides
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "brick"), (3, "brick"), (2, "shell")])
def test_mortars_across_ranks_gloo(world, kind):
    mp.spawn(_mortar_worker, args=(world, _free_port(), kind, None), nprocs=world, join=True)


def _partition_cases():
    brick = domain.Brick([0, 0, 0], [1, 1, 1], [2, 1, 2], 2)
    yield "periodic brick", brick.neighbors(), None, None, None, False
    open_brick = domain.Brick([0, 0, 0], [1, 1, 1], [1, 2, 1], 2, periodic=(False, True, False))
    yield "open brick with ghost boundary slots", open_brick.neighbors(), None, None, None, True
    shell = domain.SphericalShell(1.9, 2.9, [(1, 1), (1, 0)], 2, radial_partitioning=(2.3,))
    nd, perm = shell.neighbor_orientations()
    yield "shell", shell.neighbors(), nd, perm, None, True
    rb = domain.RefinedBrick([0, 0, 0], [1, 1, 1], [1, 1, 1], 2,
                             {(0, 0, 0): (True, True, True), (1, 1, 0): (True, False, True)})
    yield "refined brick", rb.neighbors(), None, None, rb.mortars(), False
    ws = domain.SphericalShell(1.9, 2.9, [[(1, 1), (0, 0), (0, 0), (0, 0), (1, 0), (0, 0)]], 2)
    nd, perm = ws.neighbor_orientations()
    yield "wedge-refined shell", ws.neighbors(), nd, perm, ws.mortars(), True


def test_cpp_partition_matches_python_partition():
    """DgPartition of SpectreShims.hpp (the C++ host side of the multi-GPU schedule) gives
    the same local order, local tables, ghost slots, send map and per-peer counts as
    spectre_b200.domain.Partition, for bricks, shells with non-aligned blocks, and
    aligned / oriented mortars cut across ranks."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "_build", "orientation_codes")
    src = os.path.join(root, "tests", "helpers", "orientation_codes.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-o", exe, src, "-L",
                           os.path.join(root, "spectre_b200"), "-ldgrhs",
                           "-Wl,-rpath," + os.path.join(root, "spectre_b200")])
    checked = 0
    for name, nbr, nd, perm, mt, slots in _partition_cases():
        ne = nbr.shape[0]
        mt_arr = np.zeros((0, 6), dtype=np.int64) if mt is None else np.asarray(mt)
        for world in (1, 2, 3):
            for rank in range(world):
                want = domain.Partition(nbr, world, rank, boundary_slots=slots, neighbor_direction=nd,
                                        face_permutation=perm, mortars=mt)
                text = [f"P {ne} {world} {rank} {int(nd is not None)} {int(slots)} {len(mt_arr)}",
                        " ".join(map(str, nbr.reshape(-1).tolist()))]
                if nd is not None:
                    text += [" ".join(map(str, np.asarray(nd).reshape(-1).tolist())),
                             " ".join(map(str, np.asarray(perm).reshape(-1).tolist()))]
                text.append(" ".join(map(str, mt_arr.reshape(-1).tolist())))
                out = subprocess.run([exe], input="\n".join(text) + "\n", capture_output=True,
                                     text=True)
                assert out.returncode == 0, (name, out.stderr)
                got = {ln.split()[0]: np.array(ln.split()[1:], dtype=np.int64)
                       for ln in out.stdout.strip().splitlines()}
                tag = (name, world, rank)
                assert got["counts"].tolist() == [want.n_local, want.n_interior, want.n_recv,
                                                  want.n_ghost], tag
                np.testing.assert_array_equal(got["global_ids"], want.global_ids, err_msg=str(tag))
                np.testing.assert_array_equal(got["neighbors"].reshape(-1, 6), want.local_neighbors,
                                              err_msg=str(tag))
                np.testing.assert_array_equal(got["directions"].reshape(-1, 6),
                                              want.local_neighbor_direction, err_msg=str(tag))
                np.testing.assert_array_equal(got["permutations"].reshape(-1, 6),
                                              want.local_face_permutation, err_msg=str(tag))
                np.testing.assert_array_equal(got["mortars"].reshape(-1, 6), want.local_mortars,
                                              err_msg=str(tag))
                np.testing.assert_array_equal(got["send_map"].reshape(-1, 2), want.send_map,
                                              err_msg=str(tag))
                assert got["send_counts"].tolist() == list(want.send_counts), tag
                assert got["recv_counts"].tolist() == list(want.recv_counts), tag
                assert got["external_faces"].reshape(-1, 3).tolist() == \
                    [list(t) for t in want.external_faces], tag
                checked += 1
    assert checked == 5 * 6
