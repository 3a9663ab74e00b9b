"""The N>1 path on CPU: the library's decomposition (isca_b200_decomposition) and the two Fourier
buffer layouts the CUDA kernels index into, exercised with a world_size-2 gloo process group.

Each rank runs the oracle's Legendre step for ITS zonal wavenumbers, writes the m-owner layout
([dest][mi][lat_loc][C]), exchanges the per-peer slabs (the NCCL grouped send/recv of core.cu:
exchange_fourier, here gloo isend/irecv), reads the lat-owner layout ([pos[m]][lat_loc][C]) and runs
the FFT for ITS latitudes; the gathered grid must equal the single-process oracle transform.  The
reverse direction is checked the same way.  No GPU is used."""
import os
import socket
import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_decomposition_partitions_and_balances(lib_built):
    from isca_b200 import api
    for res in ((64, 32, 21), (256, 128, 85), (512, 256, 170), (1024, 512, 341)):
        cfg = api.make_config(lon_max=res[0], lat_max=res[1], num_fourier=res[2], num_spherical=res[2] + 1)
        M = res[2]
        for P in (1, 2, 4, 8):
            decs = [api.decomposition(cfg, r, P) for r in range(P)]
            lats = np.concatenate([np.arange(d["lat_start"], d["lat_start"] + d["lat_count"]) for d in decs])
            assert np.array_equal(lats, np.arange(res[1]))                       # contiguous equal blocks (spec_mpp.F90:61-75)
            ms = np.sort(np.concatenate([d["m_list"] for d in decs]))
            assert np.array_equal(ms, np.arange(M + 1))                          # every m owned exactly once
            for r, d in enumerate(decs):
                assert np.all(d["owner"][d["m_list"]] == r)
                assert np.all(np.diff(d["m_list"]) > 0)
            assert np.array_equal(np.sort(decs[0]["pos"]), np.arange(M + 1))     # pos is a permutation
            for d in decs[1:]:
                assert np.array_equal(d["pos"], decs[0]["pos"]) and np.array_equal(d["owner"], decs[0]["owner"])
            rows = [int(np.sum(M - d["m_list"] + 2)) for d in decs]              # triangle rows per rank
            assert max(rows) <= 1.05 * min(rows) + 2 * (M + 2) / P + 8           # balanced (snake order)
    with pytest.raises(api.IscaError):
        api.decomposition(api.make_config(lat_max=64), 0, 3)                     # P must divide lat_max


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from isca_b200 import api
    from oracle.isca_oracle import Tables, Transforms, held_suarez_config
    cfg = held_suarez_config("T21", 4, 1200.0)
    tb = Tables(cfg); tr = Transforms(tb)
    I, J, M, N = cfg.lon_max, cfg.lat_max, cfg.num_fourier, cfg.num_spherical
    dec = api.decomposition(api.config_from_namelist_object(cfg), rank, world)
    m_list, pos, owner = dec["m_list"], dec["pos"], dec["owner"]
    nm, Jloc, j0 = len(m_list), dec["lat_count"], dec["lat_start"]
    nm_rank = [int(np.sum(owner == r)) for r in range(world)]
    roff = np.concatenate([[0], np.cumsum(nm_rank)])
    L = 3                                                     # levels in the batch; C = 2*L doubles per (m, lat)
    rng = np.random.default_rng(7)
    spec = (rng.standard_normal((L, N + 1, M + 1)) + 1j * rng.standard_normal((L, N + 1, M + 1))) * tb.triangle_mask
    spec[:, :, 0] = spec[:, :, 0].real

    def exchange(send_blocks, recv_shapes):
        recv = [torch.zeros(s, dtype=torch.complex128) for s in recv_shapes]
        reqs = []
        for r in range(world):
            if r == rank:
                recv[r].copy_(torch.from_numpy(send_blocks[r]))
            else:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send_blocks[r])), r))
                reqs.append(dist.irecv(recv[r], r))
        for q in reqs:
            q.wait()
        return [x.numpy() for x in recv]

    # ---- inverse: Legendre for my m's -> layout A -> exchange -> layout B -> FFT for my latitudes
    four_all = tr.spherical_to_fourier(spec)                               # [L, J, M+1] (oracle, all m) -- I only use my columns
    A = np.zeros((world, nm, Jloc, L), dtype=np.complex128)                # [(s*nm + mi)*Jloc + jl][C]
    for mi, m in enumerate(m_list):
        for j in range(J):
            s, jl = divmod(j, Jloc)
            A[s, mi, jl, :] = four_all[:, j, m]
    got = exchange([A[s] for s in range(world)], [(nm_rank[r], Jloc, L) for r in range(world)])
    B = np.zeros((M + 1, Jloc, L), dtype=np.complex128)                    # [pos[m]*Jloc + jl][C]
    for r in range(world):
        B[roff[r]:roff[r] + nm_rank[r]] = got[r]
    four_loc = np.zeros((L, Jloc, M + 1), dtype=np.complex128)
    for m in range(M + 1):
        four_loc[:, :, m] = B[pos[m]].T
    grid_loc = tr.fourier_to_grid(four_loc)                                # my latitude block
    ref = tr.spherical_to_grid(spec)
    err_inv = np.abs(grid_loc - ref[:, j0:j0 + Jloc, :]).max()

    # ---- forward: FFT for my latitudes -> layout B -> exchange -> layout A -> Legendre for my m's
    f_loc = tr.grid_to_fourier(ref[:, j0:j0 + Jloc, :])                    # [L, Jloc, M+1]
    Bf = np.zeros((M + 1, Jloc, L), dtype=np.complex128)
    for m in range(M + 1):
        Bf[pos[m]] = f_loc[:, :, m].T
    got = exchange([Bf[roff[r]:roff[r] + nm_rank[r]] for r in range(world)], [(nm, Jloc, L) for _ in range(world)])
    Af = np.stack(got)                                                     # [src s][mi][jl][C]
    four_m = np.zeros((L, J, M + 1), dtype=np.complex128)
    for mi, m in enumerate(m_list):
        for j in range(J):
            s, jl = divmod(j, Jloc)
            four_m[:, j, m] = Af[s, mi, jl, :]
    spec_back = tr.fourier_to_spherical(four_m) * tb.triangle_mask
    err_fwd = np.abs(spec_back[:, :, m_list] - spec[:, :, m_list]).max()
    np.save(os.path.join(out_dir, f"err_{rank}.npy"), np.array([err_inv, err_fwd]))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_fourier_transpose(lib_built, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        e = np.load(tmp_path / f"err_{r}.npy")
        assert e[0] < 1e-12 and e[1] < 1e-12, (r, e)


# ---------------------------------------------------------------------------------------------
# latitude halo of the grid tracer step (tracer.cu: tracer_halo_pack_kernel, core.cu: exchange_tracer_halo)
# ---------------------------------------------------------------------------------------------
def _halo_worker(rank, world, port, out_dir):
    """Each rank owns a latitude block of (tr0, u, v), packs its 2 edge rows per neighbour in the library's [3][K][2][I] layout,
    exchanges them (gloo isend/irecv in place of the grouped ncclSend/ncclRecv) and advects with ONLY its block + halos known
    (everything else zeroed); its rows must equal the single-process oracle result: one 2-row exchange of (tr0, u, v) is
    enough, q1 = q + semi_x(q) being recomputed on the halo rows (the reference exchanges three times, fv_advection.F90:161-196)."""
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from isca_b200 import api
    from oracle.isca_oracle import Tables, held_suarez_config
    from oracle.fv_advection import FVGrid, a_grid_horiz_advection
    cfg = held_suarez_config("T21", 3, 1200.0, num_tracers=1)
    tb = Tables(cfg)
    g = FVGrid(cfg, tb)
    I, J, K = cfg.lon_max, cfg.lat_max, cfg.num_levels
    dec = api.decomposition(api.config_from_namelist_object(cfg), rank, world)
    Jloc, j0 = dec["lat_count"], dec["lat_start"]
    rng = np.random.default_rng(11)                           # the same global fields on every rank
    q = rng.random((K, J, I))
    u = 40.0 * rng.standard_normal((K, J, I))                 # Courant numbers above 1 near the poles: integer fluxes too
    v = 10.0 * rng.standard_normal((K, J, I))
    dt = 1200.0
    ref = a_grid_horiz_advection(g, u, v, q, dt, np.zeros_like(q))
    mine = [f[:, j0:j0 + Jloc, :] for f in (q, u, v)]
    send_s = np.stack([f[:, 0:2, :] for f in mine])           # [3][K][2][I]: local rows 0, 1
    send_n = np.stack([f[:, Jloc - 2:Jloc, :] for f in mine])  # local rows Jloc-2, Jloc-1
    halo_s, halo_n = torch.zeros(3, K, 2, I, dtype=torch.float64), torch.zeros(3, K, 2, I, dtype=torch.float64)
    reqs = []
    if rank > 0:
        reqs += [dist.isend(torch.from_numpy(np.ascontiguousarray(send_s)), rank - 1), dist.irecv(halo_s, rank - 1)]
    if rank < world - 1:
        reqs += [dist.isend(torch.from_numpy(np.ascontiguousarray(send_n)), rank + 1), dist.irecv(halo_n, rank + 1)]
    for r in reqs:
        r.wait()
    known = [np.zeros((K, J, I)) for _ in range(3)]
    for f in range(3):
        known[f][:, j0:j0 + Jloc, :] = mine[f]
        if rank > 0:
            known[f][:, j0 - 2:j0, :] = halo_s[f].numpy()
        if rank < world - 1:
            known[f][:, j0 + Jloc:j0 + Jloc + 2, :] = halo_n[f].numpy()
    got = a_grid_horiz_advection(g, known[1], known[2], known[0], dt, np.zeros_like(q))
    err = np.abs(got[:, j0:j0 + Jloc, :] - ref[:, j0:j0 + Jloc, :]).max()
    np.save(os.path.join(out_dir, f"halo_err_{rank}.npy"), np.array([err]))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_tracer_halo(lib_built, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_halo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(tmp_path / f"halo_err_{r}.npy")[0] == 0.0, r
