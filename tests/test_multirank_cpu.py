"""CPU, world_size 2 over gloo: the host-side logic of the z-slab (N > 1) path.

The product's message list for one halo exchange (ppk_mhd3d_halo_plan, the list the engine posts inside one NCCL
group) is driven here over torch.distributed/gloo on numpy arrays, with the plain-C oracle doing the arithmetic of
each slab.  Two slabs advanced this way must be bit-identical to the undecomposed oracle run: this pins the plan's
offsets/peers, the "x,y ghosts first, then z planes" order (SolverBase.cpp:618-691) and the global-min dt
(SolverBase.cpp:152-165) without a GPU.
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

OT = "[OrszagTang]\nkt=0.5\n"  # z-period = the GLOBAL box (kt=1 would make every 8-cell slab periodic by itself)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _exchange(plan, U):
    """post the plan's messages over gloo (same order on every rank, like the NCCL group)"""
    flat = torch.from_numpy(U.reshape(-1))
    ops, recvs = [], []
    for peer, is_send, var, off, cnt in plan:
        if is_send:
            ops.append(dist.isend(flat[off:off + cnt].clone(), peer))
        else:
            buf = torch.empty(cnt, dtype=torch.float64)
            recvs.append((off, cnt, buf))
            ops.append(dist.irecv(buf, peer))
    for o in ops:
        o.wait()
    for off, cnt, buf in recvs:
        flat[off:off + cnt] = buf


def _exchange_face(plan, U, direction):
    """the packed x / y exchange of a block decomposition: pack (the layout include/ppkmhd_b200.h documents for
    ppk_face_msg = the reference's border buffers), post the plan over gloo, unpack"""
    nv, ks, js, isz = U.shape
    ops, recvs = [], []
    for peer, is_send, hi_face, first, cnt in plan:
        sl = (slice(None), slice(None), slice(None), slice(first, first + 3)) if direction == 0 else (slice(None), slice(None), slice(first, first + 3), slice(None))
        # C order of U[v, k, j, i] restricted to 3 layers == buf[g + gw*(j + jsize*(k + ksize*v))] (x) / buf[i + isize*(g + gw*(k + ksize*v))] (y)
        assert cnt == U[sl].size
        if is_send:
            ops.append(dist.isend(torch.from_numpy(np.ascontiguousarray(U[sl]).reshape(-1)), peer))
        else:
            buf = torch.empty(cnt, dtype=torch.float64)
            recvs.append((sl, buf))
            ops.append(dist.irecv(buf, peer))
    for o in ops:
        o.wait()
    for sl, buf in recvs:
        U[sl] = buf.numpy().reshape(U[sl].shape)


def _block_worker(rank, world, port, ini, nsteps, out_dir):
    """one sub-domain of an (mx, my, mz) block decomposition, arithmetic by the oracle, exchanges by the product's plans"""
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ppkmhd_b200 as ppk
    from oracle import oracle as O

    L = O.lib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    p, t_end, _ = ppk.params_from_ini(ini, rank_z=rank)
    assert rank == (p.rank_x * p.my + p.rank_y) * p.mz + p.rank_z
    plans = [ppk.face_plan(p, 0), ppk.face_plan(p, 1)]
    zplan = ppk.halo_plan(p)
    orc = O.Oracle(ini, rank_pos=(p.rank_x, p.rank_y, p.rank_z))
    assert np.array_equal(ppk.init_condition_from_ini(ini, rank_z=rank), orc.U)
    U, U2, Q = orc.U, orc.U2, orc.Q
    t = 0.0
    for _ in range(nsteps):
        for d in range(2):                      # X then Y: physical faces locally, the others through the packed exchange
            got = {hi for _, is_send, hi, _, _ in plans[d] if not is_send}
            for side in range(2):
                if side not in got:
                    L.orc_make_boundary(C.byref(orc.p), dp(U), 2 * d + side)
            _exchange_face(plans[d], U, d)
        _exchange(zplan, U)
        ncell = U[0].size
        recv_k0 = {(off % ncell) // (U.shape[2] * U.shape[3]) for _, is_send, _, off, _ in zplan if not is_send}
        for f, k0 in ((4, 0), (5, p.nz + 3)):
            if k0 not in recv_k0:
                L.orc_make_boundary(C.byref(orc.p), dp(U), f)
        L.orc_convert_to_primitives(C.byref(orc.p), dp(U), dp(Q))
        inv = torch.tensor([L.orc_compute_inv_dt(C.byref(orc.p), dp(Q))], dtype=torch.float64)
        dist.all_reduce(inv, op=dist.ReduceOp.MAX)
        dt = orc.p.cfl / float(inv[0])
        L.orc_godunov_v0(C.byref(orc.p), dp(U), dp(Q), dp(U2), orc.scratch, dt)
        U, U2 = U2, U
        t += dt
    np.save(os.path.join(out_dir, f"block{rank}.npy"), U[:, 3:-3, 3:-3, 3:-3])
    np.save(os.path.join(out_dir, f"t{rank}.npy"), np.array([t]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,bc", [((2, 2, 1), 3), ((1, 2, 2), 3), ((2, 1, 2), [1, 2, 3, 3, 2, 1])])
def test_blocks_over_gloo_match_undecomposed_oracle(tmp_path, oracle_mod, shape, bc):
    """Block (pencil) decomposition, world_size 4 on CPU: the product's face plans (ppk_mhd3d_face_plan: peers, layers, packed
    layout) and its rank <-> (x, y, z) mapping driven over gloo with the oracle doing the arithmetic must reproduce the
    undecomposed oracle run bit for bit -- edges and corners included, which only the X -> Y -> Z order gets right."""
    O = oracle_mod
    mx, my, mz = shape
    nsteps = 2
    n = (12, 10, 8)
    kw = dict(nstepmax=nsteps, extra=OT, tend=10.0, bc=bc)
    ini_n = O.make_ini("orszag_tang", n, mx=mx, my=my, mz=mz, **kw)
    ini_1 = O.make_ini("orszag_tang", (n[0] * mx, n[1] * my, n[2] * mz), **kw)
    port = _free_port()
    world = mx * my * mz
    mp.spawn(_block_worker, args=(world, port, ini_n, nsteps, str(tmp_path)), nprocs=world, join=True)
    whole = O.Oracle(ini_1).run(nsteps)
    got = np.empty_like(whole.interior())
    for r in range(world):
        cz, cy, cx = r % mz, (r // mz) % my, r // (mz * my)
        got[:, cz * n[2]:(cz + 1) * n[2], cy * n[1]:(cy + 1) * n[1], cx * n[0]:(cx + 1) * n[0]] = np.load(tmp_path / f"block{r}.npy")
        assert float(np.load(tmp_path / f"t{r}.npy")[0]) == whole.t
    assert np.array_equal(got, whole.interior()), "block-decomposed run differs from the undecomposed run"


def _worker(rank, world, port, ini, nsteps, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ppkmhd_b200 as ppk
    from oracle import oracle as O

    L = O.lib()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    p, t_end, _ = ppk.params_from_ini(ini, rank_z=rank)
    plan = ppk.halo_plan(p)
    ncell = (p.nx + 6) * (p.ny + 6) * (p.nz + 6)
    recv_k0 = {(off % ncell) // ((p.nx + 6) * (p.ny + 6)) for _, is_send, _, off, _ in plan if not is_send}
    # z faces the exchange fills (inner faces, and outer faces of a periodic box: SolverBase.cpp:672-689)
    exchanged = {4: 0 in recv_k0, 5: (p.nz + 3) in recv_k0}
    orc = O.Oracle(ini, rank_pos=(0, 0, rank))
    # the product's host layer and the oracle agree on the slab's initial array
    assert np.array_equal(ppk.init_condition_from_ini(ini, rank_z=rank), orc.U)
    U, U2, Q = orc.U, orc.U2, orc.Q
    t = 0.0
    for _ in range(nsteps):
        for f in range(4):                      # x then y faces (physical or periodic), full transverse extent
            L.orc_make_boundary(C.byref(orc.p), dp(U), f)
        _exchange(plan, U)                      # z planes carry the already-filled x/y ghosts
        for f in (4, 5):                        # z faces with a physical BC (not filled by the exchange)
            if not exchanged[f]:
                L.orc_make_boundary(C.byref(orc.p), dp(U), f)
        L.orc_convert_to_primitives(C.byref(orc.p), dp(U), dp(Q))
        inv = torch.tensor([L.orc_compute_inv_dt(C.byref(orc.p), dp(Q))], dtype=torch.float64)
        dist.all_reduce(inv, op=dist.ReduceOp.MAX)   # engine: ncclAllReduce(max) of 1/dt == MIN of dt
        dt = orc.p.cfl / float(inv[0])
        if t + dt > t_end:
            dt = t_end - t
        L.orc_godunov_v0(C.byref(orc.p), dp(U), dp(Q), dp(U2), orc.scratch, dt)
        U, U2 = U2, U
        t += dt
    np.save(os.path.join(out_dir, f"slab{rank}.npy"), U[:, 3:-3, 3:-3, 3:-3])
    np.save(os.path.join(out_dir, f"t{rank}.npy"), np.array([t]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("bcz", [3, 2])
def test_two_slabs_over_gloo_match_undecomposed_oracle(tmp_path, oracle_mod, bcz):
    O = oracle_mod
    nsteps = 3
    kw = dict(nstepmax=nsteps, extra=OT, tend=10.0)
    ini2 = O.make_ini("orszag_tang", (16, 12, 8), mz=2, **kw)
    ini1 = O.make_ini("orszag_tang", (16, 12, 16), **kw)
    if bcz != 3:  # physical (Neumann) z faces on the outer slabs: only the inner face is exchanged
        ini2 = ini2.replace("boundary_type_zmin=3", f"boundary_type_zmin={bcz}").replace("boundary_type_zmax=3", f"boundary_type_zmax={bcz}")
        ini1 = ini1.replace("boundary_type_zmin=3", f"boundary_type_zmin={bcz}").replace("boundary_type_zmax=3", f"boundary_type_zmax={bcz}")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ini2, nsteps, str(tmp_path)), nprocs=2, join=True)
    whole = O.Oracle(ini1).run(nsteps)
    got = np.concatenate([np.load(tmp_path / "slab0.npy"), np.load(tmp_path / "slab1.npy")], axis=1)
    assert float(np.load(tmp_path / "t0.npy")[0]) == whole.t == float(np.load(tmp_path / "t1.npy")[0])
    assert np.array_equal(got, whole.interior()), "two z-slabs differ from the undecomposed run"


def test_halo_plan_shapes():
    import ppkmhd_b200 as ppk
    from oracle import oracle as O

    ini = O.make_ini("orszag_tang", (16, 12, 8), mz=4, extra=OT)
    for r in range(4):
        p, _, _ = ppk.params_from_ini(ini, rank_z=r)
        plan = ppk.halo_plan(p)
        assert len(plan) == 32                                  # 8 variables x (2 sends + 2 receives), periodic ring
        plane, ncell = 22 * 18, 22 * 18 * 14
        for peer, is_send, var, off, cnt in plan:
            assert cnt == 3 * plane and peer in ((r - 1) % 4, (r + 1) % 4)
            k0 = (off - var * ncell) // plane
            assert (off - var * ncell) % plane == 0
            assert k0 in ((3, 8) if is_send else (0, 11))
    # non-periodic: outer faces are physical => the end slabs exchange one face only
    ini_d = ini.replace("boundary_type_zmin=3", "boundary_type_zmin=1").replace("boundary_type_zmax=3", "boundary_type_zmax=1")
    assert [len(ppk.halo_plan(ppk.params_from_ini(ini_d, rank_z=r)[0])) for r in range(4)] == [16, 32, 32, 16]
    p1, _, _ = ppk.params_from_ini(O.make_ini("orszag_tang", (16, 12, 8), extra=OT))
    assert ppk.halo_plan(p1) == []
