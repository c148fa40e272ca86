"""Band partition of ONE multi-band job over the GPUs of a box, with plane offload for the heavy bands.

The reference runs one actor per imaging band (``/root/reference/src/pfb_imaging/operators/band_worker.py:217-246``)
and the bands never talk to each other on the measurement-operator path, so a job of ``nband`` bands shards over
``world`` GPUs with no data-path collective.  Two things limit that partition:

* bands are not equally expensive (C2: band 7 costs 1.7x band 0 — more w-planes, a larger active uv window), so
  the bands are assigned longest-processing-time first (:func:`lpt_assign`);
* with one band per GPU the heaviest band bounds the step (C2: sum / max = 6.3 on 8 GPUs).  The plane transforms are
  the divisible half of a Hessian apply: :func:`plan_offloads` moves the last ``nq`` w-planes of a heavy band to the
  least loaded GPU, whose transform kernels read / write the owner's plane stack over NVLink (``pfbg_split_*`` in
  ``include/pfbgrid.h``; row additivity of the operator: ``tests/test_imager_pass2.py:45-63``).

Host logic only (lists and dicts); the device work is in ``libpfbgrid.so``.
"""

from __future__ import annotations

import os

# A helper's transform kernels spin on flags in device memory while the same GPU runs its own band on other streams:
# give every stream its own hardware queue (the default of 8 connections aliases streams, and the band's kernels would
# line up behind a waiting helper kernel).  Only effective when set before the CUDA context is created.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .wgridder import GridderPlan, SplitHelper  # noqa: E402

# cost model of one offloaded plane, relative to the owner's measured per-plane transform time: the helper's column
# passes run against peer memory (NVLink sectors instead of L2 hits), and the first plane of an offload also pays
# for the image copy, the partial-image copy and the flag waits
HELPER_PLANE_FACTOR = 1.35
HELPER_FIXED_MS = 0.45
OWNER_FIXED_MS = 0.15
MAX_OFFLOAD_PLANES = 6


def lpt_assign(costs, world):
    """Longest-processing-time-first partition: returns the owning rank of every band."""
    order = sorted(range(len(costs)), key=lambda b: (-costs[b], b))
    load = [0.0] * world
    owner = [0] * len(costs)
    for b in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[b] = r
        load[r] += costs[b]
    return owner


def plan_offloads(costs, t_plane, nplanes, owner, world, max_planes=MAX_OFFLOAD_PLANES,
                  helper_factor=HELPER_PLANE_FACTOR, helper_fixed=HELPER_FIXED_MS, owner_fixed=OWNER_FIXED_MS):
    """Greedy plane offload.  costs[b]: ms per apply of band b alone; t_plane[b]: ms of forward + inverse transform
    work per plane of band b; nplanes[b]; owner[b]: rank.  Returns (offloads, loads): offloads maps band ->
    dict(owner, helper, nq); loads are the modelled per-rank ms after the moves.  Deterministic: every rank
    computes the same schedule from the same (all-gathered) inputs."""
    nb = len(costs)
    load = [0.0] * world
    for b in range(nb):
        load[owner[b]] += costs[b]
    off = {}
    if world < 2:
        return off, load
    for _ in range(nb * max_planes):
        r = max(range(world), key=lambda k: (load[k], -k))
        best = None
        for b in range(nb):
            if owner[b] != r:
                continue
            cur = off.get(b)
            if cur is not None:
                if cur["nq"] >= min(max_planes, nplanes[b] - 1):
                    continue
                h, first = cur["helper"], False
            else:
                if nplanes[b] < 2:
                    continue
                h, first = min((k for k in range(world) if k != r), key=lambda k: (load[k], k)), True
            new_r = load[r] - t_plane[b] + (owner_fixed if first else 0.0)
            new_h = load[h] + t_plane[b] * helper_factor + (helper_fixed if first else 0.0)
            new_max = max(new_r, new_h)
            if best is None or new_max < best[0]:
                best = (new_max, b, h, first, new_r, new_h)
        if best is None or best[0] >= load[r] - 1e-6:
            break
        _, b, h, first, new_r, new_h = best
        load[r], load[h] = new_r, new_h
        if first:
            off[b] = dict(owner=r, helper=h, nq=1)
        else:
            off[b]["nq"] += 1
    return off, load


class BandSplit:
    """Set-up and per-step service of the offloads this process takes part in.

    `plans`: band -> bound :class:`GridderPlan` of the bands this rank owns; `offloads`: output of
    :func:`plan_offloads`; `gather(obj)` returns the list of every rank's `obj` (``dist.all_gather_object`` in a
    job, ``lambda o: [o]`` inside one process, where a rank may be owner and helper at once)."""

    def __init__(self, plans: dict, offloads: dict, rank: int, device: int, gather):
        self.rank, self.device = rank, device
        self.owned = {b: plans[b] for b, o in offloads.items() if o["owner"] == rank}
        self.helpers: dict[int, SplitHelper] = {}
        offers = {}
        for b, gp in self.owned.items():
            assert isinstance(gp, GridderPlan)
            nq = offloads[b]["nq"]
            offers[b] = dict(plan=gp.plan, window=gp.window(), blobs=gp.split_owner_init(nq), nq=nq)
        all_offers = {}
        for d in gather(offers):
            all_offers.update(d)
        replies = {}
        for b, o in sorted(offloads.items()):
            if o["helper"] != rank:
                continue
            m = all_offers[b]
            self.helpers[b] = SplitHelper(m["plan"], m["nq"], m["window"], m["blobs"], device=device)
            replies[b] = self.helpers[b].mailbox_blob
        all_replies = {}
        for d in gather(replies):
            all_replies.update(d)
        for b, gp in self.owned.items():
            gp.split_owner_connect(all_replies[b])

    def serve(self, streams=None):
        """Enqueue one apply worth of helper work for every band this rank helps (asynchronous).  `streams`: band ->
        cudaStream_t handle (int) or one handle for all."""
        for b, h in self.helpers.items():
            st = streams.get(b) if isinstance(streams, dict) else streams
            h.serve(st)

    def check(self):
        """Synchronise and raise if a flag wait gave up (owner and helper out of step, or the peer is gone)."""
        bad = [b for b, gp in self.owned.items() if gp.split_timed_out()] + \
              [b for b, h in self.helpers.items() if h.timed_out()]
        if bad:
            raise RuntimeError(f"band split: flag wait timed out for band(s) {sorted(set(bad))}")

    def close(self):
        """Both sides must be idle (synchronise + barrier first)."""
        for gp in self.owned.values():
            gp.split_end()
        for h in self.helpers.values():
            h.close()
        self.helpers.clear()
        self.owned = {}
