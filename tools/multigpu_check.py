"""Run under torch.distributed.run with N ranks: the N-rank run must reproduce the 1-rank run -- Held-Suarez core with the sphum
grid tracer (latitude halo exchange) and the idealized moist model.  (test infrastructure)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from isca_b200 import api, moist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = sys.argv[1] if len(sys.argv) > 1 else "T42"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    moist_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    import bench

    def new_uid():
        box = [api.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def map_peers(core):
        if os.environ.get("ISCA_B200_NO_P2P") is None:
            handles = [None] * world
            dist.all_gather_object(handles, core.ipc_handles())
            core.set_peer_handles(handles)

    import json
    report = {}
    nml = bench.hs_namelist(res, K, True)
    nml["initial_sphum"] = 1.0e-3                      # a non-trivial tracer field from the start
    cfg = api.make_config(**nml)
    atm = api.Atmosphere(cfg, rank=rank, nranks=world, nccl_unique_id=new_uid())
    map_peers(atm)
    atm.cold_start()
    atm.atmosphere(nsteps)

    def snap(a):
        return dict(u=a.get_field(api.F_U), v=a.get_field(api.F_V), T=a.get_field(api.F_T), ps=a.get_field(api.F_PS),
                    vor=a.get_field(api.F_VOR), div=a.get_field(api.F_DIV), q=a.get_field(api.F_TRACER0),
                    vors=a.get_spectral(api.S_VOR), divs=a.get_spectral(api.S_DIV), ts=a.get_spectral(api.S_T),
                    lnps=a.get_spectral(api.S_LNPS), vors_prev=a.get_spectral(api.S_VOR, api.LEVEL_PREVIOUS),
                    divs_prev=a.get_spectral(api.S_DIV, api.LEVEL_PREVIOUS), ts_prev=a.get_spectral(api.S_T, api.LEVEL_PREVIOUS),
                    lnps_prev=a.get_spectral(api.S_LNPS, api.LEVEL_PREVIOUS))
    loc = snap(atm)
    torch.cuda.synchronize()
    t0 = time.time(); atm.atmosphere(50); torch.cuda.synchronize(); dt_ms = (time.time() - t0) / 50 * 1e3
    ms = atm.get_scalar(api.SC_LAST_STEP_MS)
    prof = atm.profile_step(10)
    gathered = [None] * world
    dist.all_gather_object(gathered, loc)
    atm.atmosphere_end()

    mloc = None
    if moist_steps > 0:
        mm = moist.frierson_test_case(res, K, bench.RES[res][3], rank=rank, nranks=world, nccl_unique_id=new_uid())
        map_peers(mm.core)
        mm.core.cold_start(); mm.idealized_moist_phys_init(); mm.atmosphere(moist_steps)
        mloc = dict(T=mm.core.get_field(api.F_T), q=mm.core.get_field(api.F_TRACER0), u=mm.core.get_field(api.F_U),
                    t_surf=mm.get("t_surf"), precip=mm.get("precip"), flux_q=mm.get("flux_q"))
        mm.atmosphere(30)
        mms = mm.timing()
        mm.atmosphere_end()
    mg = [None] * world
    dist.all_gather_object(mg, mloc)

    if rank == 0:
        one = api.Atmosphere(cfg)
        one.cold_start(); one.atmosphere(nsteps)
        ref = snap(one)
        one.atmosphere(50)
        ms1 = one.get_scalar(api.SC_LAST_STEP_MS)
        one.atmosphere_end()
        out = {}
        for k in ("u", "v", "T", "ps", "vor", "div", "q"):
            full = np.concatenate([g[k] for g in gathered], axis=-2)
            out[k] = float(np.abs(full - ref[k]).max() / np.abs(ref[k]).max())
        for k in ("vors", "divs", "ts", "lnps", "vors_prev", "divs_prev", "ts_prev", "lnps_prev"):
            full = sum(g[k] for g in gathered)               # every rank returns zeros for the m it does not own
            out[k] = float(np.abs(full - ref[k]).max() / np.abs(ref[k]).max())
        report["dry"] = out
        report["max_abs"] = {k: float(np.abs(ref[k]).max()) for k in ("vors", "divs", "u", "v")}
        print(f"MULTIGPU {res} L{K} P={world} (HS + sphum tracer): rel diff vs 1 rank after {nsteps} steps: {out}")
        print(f"MULTIGPU ms/step P={world}: {ms:.3f} (events) {dt_ms:.3f} (wall); 1 rank: {ms1:.3f}")
        print("MULTIGPU groups:", {k: round(v, 4) for k, v in prof.items()})
        if moist_steps > 0:
            m1 = moist.frierson_test_case(res, K, bench.RES[res][3])
            m1.core.cold_start(); m1.idealized_moist_phys_init(); m1.atmosphere(moist_steps)
            mref = dict(T=m1.core.get_field(api.F_T), q=m1.core.get_field(api.F_TRACER0), u=m1.core.get_field(api.F_U),
                        t_surf=m1.get("t_surf"), precip=m1.get("precip"), flux_q=m1.get("flux_q"))
            m1.atmosphere(30)
            mms1 = m1.timing()
            m1.atmosphere_end()
            mo = {}
            for k in mref:
                full = np.concatenate([g[k] for g in mg], axis=-2)
                mo[k] = float(np.abs(full - mref[k]).max() / max(np.abs(mref[k]).max(), 1e-300))
            report["moist"] = mo
            print(f"MULTIGPU moist {res} L{K} P={world}: rel diff vs 1 rank after {moist_steps} steps: {mo}")
            print(f"MULTIGPU moist ms/step P={world}: {mms[0]:.3f} (physics {mms[1]:.3f}); 1 rank: {mms1[0]:.3f} (physics {mms1[1]:.3f})")
    if rank == 0:
        print("MULTIGPU_JSON " + json.dumps(dict(res=res, K=K, ranks=world, steps=nsteps, moist_steps=moist_steps, **report)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
