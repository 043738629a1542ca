"""Calibration ensembles across GPUs -- the replacement of the reference's ``multiprocessing.Pool`` evaluation.

Reference: ``SimuBasin.Calib_NSGAII`` (wmf.py:4662-4733) evaluates every NSGA-II generation with
``__ejec_parallel__`` (wmf.py:872-878): ``Pool(nproc).map`` of ``run_shia`` over the population, keeping
``Qsim[nodo]`` of each member, then scores it with ``__eval_nash__`` / ``__eval_q_pico__`` (wmf.py:749-754, 775-781).

Here: one process per GPU (``torch.distributed``); member ``i`` of the population runs on rank ``i mod world``; every
rank holds a full replica of topology, maps and rain (members are independent, so there is no data-path collective);
the members of a rank run ``members_per_launch`` at a time in ONE wavefront launch sequence (member-minor state layout,
``wmf_shia_run`` with ``n_members > 1``); the per-member objective scores are computed on the device
(``wmf_eval_scores``) and exchanged with a single all-gather of ``[members, 2]`` fp32 -- the only collective.
"""
from __future__ import annotations

import numpy as np


def shard_members(n_members: int, rank: int, world: int) -> np.ndarray:
    """Indices of the population members evaluated by ``rank``: round-robin, as balanced as possible."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return np.arange(rank, n_members, world, dtype=np.int64)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def gather_scores(local_scores, n_members: int, device=None):
    """All-gather the per-member scores of every rank into population order.

    ``local_scores``: (len(shard_members(n_members, rank, world)), n_obj) torch tensor or numpy array on this rank.
    Returns a (n_members, n_obj) torch tensor (same on every rank).  Ranks may hold different member counts; the
    exchange is padded to the largest shard so that a single fixed-size ``all_gather`` (NCCL or gloo) does it."""
    import torch
    dist, rank, world = _dist()
    t = torch.as_tensor(local_scores, dtype=torch.float32)
    if device is not None:
        t = t.to(device)
    if t.dim() == 1:
        t = t[:, None]
    n_obj = t.shape[1]
    mine = shard_members(n_members, rank, world)
    if t.shape[0] != mine.size:
        raise ValueError(f"rank {rank} holds {t.shape[0]} score rows, its shard has {mine.size} members")
    if world == 1:
        return t.clone()
    per = (n_members + world - 1) // world
    pad = torch.full((per, n_obj), float("nan"), dtype=torch.float32, device=t.device)
    pad[: t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.empty((n_members, n_obj), dtype=torch.float32, device=t.device)
    for r in range(world):
        idx = shard_members(n_members, r, world)
        out[torch.as_tensor(idx, device=t.device)] = parts[r][: idx.size]
    return out


def share_rain(rain, device):
    """Device copy of a HOST rain table ``(n_records, N)`` int32 for an ensemble evaluation.

    Every rank of an ensemble reads the same table.  Instead of N ranks pulling N identical copies through the host's
    memory and their PCIe links, rank r uploads one contiguous block of ``ceil(R / world)`` records only (a single DMA
    from page-locked memory when the caller pinned the table) and one NCCL all-gather over NVLink hands every rank the
    other blocks, already in record order.  Bit-identical to a plain upload (integer copies).  A single process just
    uploads."""
    import torch
    dist, rank, world = _dist()
    h = torch.as_tensor(np.ascontiguousarray(rain, dtype=np.int32)) if not hasattr(rain, "is_cuda") else rain

    def settle(t):
        # the library launches on the context's own stream: what torch queued on its current stream must have landed
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        return t

    if world == 1:
        return settle(h.to(device, non_blocking=True))
    R, N = h.shape
    per = (R + world - 1) // world
    full = torch.empty((world * per, N), dtype=torch.int32, device=device)
    lo, hi = min(R, rank * per), min(R, (rank + 1) * per)
    mine = torch.empty((per, N), dtype=torch.int32, device=device)
    if hi > lo:
        mine[: hi - lo].copy_(h[lo:hi], non_blocking=True)
    if hi - lo < per:
        mine[hi - lo:].zero_()
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(full, mine)
    else:                                                # gloo (CPU tests of the record bookkeeping)
        lst = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(lst, mine)
        full = torch.cat(lst, 0)
    return settle(full[:R])


def evaluate_population(models, calibs, qobs, n_cont, n_reg, rain, pos_evento, row: int = 0,
                        members_per_launch: int = 64, run_batch=None, stats: dict | None = None):
    """Scores ``[nash, qpico]`` of every calibration set in ``calibs`` (P,11) against ``qobs`` (n_reg), sharded over the
    ranks of the initialised process group (or run locally when there is none).  Returns a (P,2) torch tensor.

    ``rain`` may be a torch CUDA tensor (used in place) or a host array: a host table is uploaded ONCE per evaluation
    (``share_rain``: 1/world of the records per rank + an NVLink all-gather), and the module's static maps are made
    resident for the duration of the call, so that the launches of a rank share them.
    ``run_batch(calib_batch) -> (m,2) scores`` may be injected (CPU tests of the sharding / gather logic); by default
    a batch is one ``wmf_shia_run`` with ``n_members = m`` followed by ``wmf_eval_scores``, all resident on the GPU.
    ``stats`` (optional dict) accumulates ``wave_ms`` (CUDA-event time of the wavefront launches), ``waves`` and
    ``batches`` of this rank; asking for it synchronises after every batch."""
    import torch
    calibs = np.ascontiguousarray(np.asarray(calibs, dtype=np.float32))
    if calibs.ndim != 2 or calibs.shape[1] != 11:
        raise ValueError("calibs must be (P,11) (modelosv2.f90:201)")
    P = calibs.shape[0]
    _, rank, world = _dist()
    mine = shard_members(P, rank, world)
    dev = None
    drop_resident = False
    if run_batch is None:
        c = models.ctx
        dev = torch.device(f"cuda:{c.device}")
        d_qobs = torch.as_tensor(np.ascontiguousarray(qobs, np.float32), device=dev)
        if d_qobs.numel() != n_reg:
            raise ValueError("qobs must hold n_reg values")
        if int(models.sim_sediments) == 1 or int(models.sim_slides) == 1:
            raise NotImplementedError("evaluate_population scores hydrographs: run it with models.sim_sediments = "
                                      "models.sim_slides = 0 (storm scenarios with sediments go through run_scenarios)")
        ch = getattr(models, "control_h", None)
        n_conth = max(1, int(np.count_nonzero(ch))) if ch is not None else 1
        if not hasattr(rain, "is_cuda") or not rain.is_cuda:
            rain = share_rain(rain, dev)
        if not object.__getattribute__(models, "__dict__").get("_dev"):
            models.make_resident()
            drop_resident = True

        def run_batch(cb):
            m = cb.shape[0]
            # the ensemble configuration (plain hydrology, member-minor outputs) for every chunk size, one member included
            out = models._run(cb, n_cont, n_conth, n_reg, rain, pos_evento, device_out=True, want={"q"}, ensemble=True)
            sc = torch.empty((m, 2), dtype=torch.float32, device=dev)
            c.check(c.L.wmf_eval_scores(c.h, out["q"].data_ptr(), m, n_cont, n_reg, row, d_qobs.data_ptr(), sc.data_ptr()))
            if stats is not None:
                stats["wave_ms"] = stats.get("wave_ms", 0.0) + c.last_kernel_ms()
                stats["waves"] = stats.get("waves", 0) + c.last_shia_stats()[1]
                stats["batches"] = stats.get("batches", 0) + 1
            return sc

    try:
        chunks = []
        for s in range(0, mine.size, max(1, members_per_launch)):
            idx = mine[s: s + members_per_launch]
            chunks.append(torch.as_tensor(run_batch(calibs[idx]), dtype=torch.float32))
    finally:
        if drop_resident:
            models.make_resident(False)
    if chunks:
        local = torch.cat(chunks, 0)
    else:
        local = torch.empty((0, 2), dtype=torch.float32, device=dev)
    return gather_scores(local, P, device=local.device)


def run_scenario_ensemble(models, calib, pos_scenarios, n_cont, n_reg, rain, n_conth: int = 1, scenarios_per_launch: int = 8,
                          device_out: bool = False, want=None):
    """Storm scenarios as a member axis (BASELINE config 5): scenario ``s`` follows its own sequence ``pos_scenarios[s]``
    (n_reg 1-based record numbers) through ONE shared table of rain records ``rain`` (n_records, N); topology, maps and
    parameters are shared, sediments and landslides may be on.  ``scenarios_per_launch`` scenarios run in one wavefront
    launch sequence (member-minor state: the waves of a basin whose single run leaves the GPU half empty fill it), the
    scenarios of the initialised process group's ranks are sharded round robin like calibration members.

    Returns ``(indices, outputs)``: the scenario indices this rank ran and, per launch, the dict ``models._run`` returns
    (arrays with the scenario axis last).  The reference runs one process per scenario (wmf.py:872-878)."""
    pos = np.ascontiguousarray(np.asarray(pos_scenarios, dtype=np.int32))
    if pos.ndim != 2 or pos.shape[1] != n_reg:
        raise ValueError("pos_scenarios must be (n_scenarios, n_reg)")
    _, rank, world = _dist()
    mine = shard_members(pos.shape[0], rank, world)
    cal = np.ascontiguousarray(np.asarray(calib, np.float32).reshape(-1))
    outs = []
    for s in range(0, mine.size, max(1, scenarios_per_launch)):
        idx = mine[s: s + scenarios_per_launch]
        if idx.size == 1:       # a lone scenario is an ordinary single run
            outs.append(models._run(cal[None, :], n_cont, n_conth, n_reg, rain, pos[idx[0]], device_out=device_out, want=want,
                                    ensemble=True))
        else:
            outs.append(models._run(np.tile(cal[None, :], (idx.size, 1)), n_cont, n_conth, n_reg, rain, pos[idx],
                                    device_out=device_out, want=want))
    return mine, outs


def run_scenarios(setup, scenarios, run_one, concurrency: int = 3, device: int | None = None):
    """Storm-scenario ensembles on ONE GPU (BASELINE config 5): scenarios differ in rainfall (or in anything else) and are
    independent runs -- the reference would fork one process per scenario (wmf.py:872-878).  Here ``concurrency`` workers,
    each with its own ``Context`` (one CUDA stream) and its own ``ModelsModule``, take the scenarios round robin from host
    threads (ctypes releases the GIL during ``wmf_shia_run``): the wavefronts of a basin whose waves do not fill the GPU --
    short runs, register-heavy sediment / landslide variants -- then overlap on the device.  Measured on the C5 basin
    (2.09 M cells, 288 steps, SED + slides): 264 ms per scenario alone, 129 ms with three in flight.

    ``setup(models)`` loads the basin into a fresh ``ModelsModule`` (e.g. ``synth.apply_to_models`` + ``make_resident``);
    ``run_one(models, scenario)`` runs one scenario and returns its result.  Returns the results in scenario order.
    Across GPUs, shard the scenario list with ``shard_members`` exactly like calibration members."""
    import threading

    from . import _lib
    from ._models import ModelsModule
    scenarios = list(scenarios)
    k = max(1, min(int(concurrency), len(scenarios)))
    workers = []
    for _ in range(k):
        m = ModelsModule(_lib.Context(device))
        setup(m)
        workers.append(m)
    results = [None] * len(scenarios)
    errors = []

    def work(w):
        try:
            for i in range(w, len(scenarios), k):
                results[i] = run_one(workers[w], scenarios[i])
            workers[w].ctx.synchronize()
        except Exception as e:          # surfaced after the join
            errors.append(e)

    threads = [threading.Thread(target=work, args=(w,)) for w in range(k)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results
