"""Run under torch.distributed.run with N ranks: the N-rank run must reproduce the 1-rank run.
(test infrastructure)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from isca_b200 import api


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = sys.argv[1] if len(sys.argv) > 1 else "T42"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    import bench
    nml = bench.hs_namelist(res, K)
    cfg = api.make_config(**nml)
    box = [api.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    atm = api.Atmosphere(cfg, rank=rank, nranks=world, nccl_unique_id=box[0])
    if os.environ.get("ISCA_B200_NO_P2P") is None:
        handles = [None] * world
        dist.all_gather_object(handles, atm.ipc_handles())
        atm.set_peer_handles(handles)
    atm.cold_start()
    atm.atmosphere(nsteps)
    loc = dict(u=atm.get_field(api.F_U), T=atm.get_field(api.F_T), ps=atm.get_field(api.F_PS), vor=atm.get_field(api.F_VOR),
               ts=atm.get_spectral(api.S_T), lnps=atm.get_spectral(api.S_LNPS), divs=atm.get_spectral(api.S_DIV))
    torch.cuda.synchronize()
    t0 = time.time(); atm.atmosphere(50); torch.cuda.synchronize(); dt_ms = (time.time() - t0) / 50 * 1e3
    ms = atm.get_scalar(api.SC_LAST_STEP_MS)
    prof = atm.profile_step(10)
    gathered = [None] * world
    dist.all_gather_object(gathered, loc)
    atm.atmosphere_end()
    if rank == 0:
        one = api.Atmosphere(cfg)
        one.cold_start(); one.atmosphere(nsteps)
        ref = dict(u=one.get_field(api.F_U), T=one.get_field(api.F_T), ps=one.get_field(api.F_PS), vor=one.get_field(api.F_VOR),
                   ts=one.get_spectral(api.S_T), lnps=one.get_spectral(api.S_LNPS), divs=one.get_spectral(api.S_DIV))
        one.atmosphere(50)
        ms1 = one.get_scalar(api.SC_LAST_STEP_MS)
        one.atmosphere_end()
        out = {}
        for k in ("u", "T", "ps", "vor"):
            full = np.concatenate([g[k] for g in gathered], axis=-2)
            out[k] = float(np.abs(full - ref[k]).max() / np.abs(ref[k]).max())
        for k in ("ts", "lnps", "divs"):
            full = sum(g[k] for g in gathered)               # every rank returns zeros for the m it does not own
            out[k] = float(np.abs(full - ref[k]).max() / np.abs(ref[k]).max())
        print(f"MULTIGPU {res} L{K} P={world}: rel diff vs 1 rank after {nsteps} steps: {out}")
        print(f"MULTIGPU ms/step P={world}: {ms:.3f} (events) {dt_ms:.3f} (wall); 1 rank: {ms1:.3f}")
        print("MULTIGPU groups:", {k: round(v, 4) for k, v in prof.items()})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
