"""One-robot-per-GPU sharding of loop-closure batches (new; the reference is single-process: "multi-robot" there means
N ROS namespaces consumed by one global_manager, GM/src/global_manager.cpp:287-442).

Rank r = GPU r = robot r.  Every rank keeps its own keyframe clouds resident.  A candidate pair (robot a, keyframe i) <->
(robot b, keyframe j) is refined on rank a (the robot that holds the source/query keyframe); the only exchange step is an
all-gather of the candidate TARGET clouds (fixed-size slots, 120k x 3 float32 = 1.44 MB each) over NCCL/NVLink at
loop-closure time, followed by an all-gather of the 4x4 results.  No other collective exists on the path.
The functions work with any torch.distributed backend (gloo on CPU tensors in the tests, nccl on CUDA tensors)."""
import numpy as np
import torch
import torch.distributed as dist


def owner_of_pair(pair, world_size):
    """pair = (src_robot, src_kf, tgt_robot, tgt_kf) -> rank that refines it (the source robot's GPU)."""
    return pair[0] % world_size


def shard_pairs(pairs, rank, world_size):
    """Deterministic partition of a global candidate list: indices of the pairs this rank refines."""
    return [i for i, p in enumerate(pairs) if owner_of_pair(p, world_size) == rank]


def requested_targets(pairs, world_size):
    """For every rank: the sorted list of (tgt_robot, tgt_kf) clouds somebody else needs from it."""
    need = [set() for _ in range(world_size)]
    for p in pairs:
        if owner_of_pair(p, world_size) != p[2] % world_size:
            need[p[2] % world_size].add((p[2], p[3]))
    return [sorted(s) for s in need]


def allgather_slots(local_slots, async_op=False):
    """local_slots: tensor [S, N, 3] (same S, N on every rank; pad with NaN rows if a robot contributes fewer clouds).
    Returns [world, S, N, 3] on every rank.  One collective per loop-closure batch.
    async_op=True returns (out, work): the collective runs on the backend's own stream while the caller prepares the clouds
    it already holds; `work.wait()` orders the current stream after it (work is None when there is nothing to wait for)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = local_slots.unsqueeze(0)
        return (out, None) if async_op else out
    world = dist.get_world_size()
    out = torch.empty((world,) + tuple(local_slots.shape), dtype=local_slots.dtype, device=local_slots.device)
    if local_slots.is_cuda:
        work = dist.all_gather_into_tensor(out, local_slots.contiguous(), async_op=async_op)  # one ncclAllGather, no staging copy
    else:
        work = dist.all_gather(list(out.unbind(0)), local_slots.contiguous(), async_op=async_op)  # gloo: views into `out`
    return (out, work) if async_op else out


def neighbor_exchange(offer, send_to, recv_from, out=None, async_op=False):
    """Point-to-point exchange of candidate clouds: this rank sends `offer` [S, N, 3] to rank `send_to` and receives the same-shaped
    block of rank `recv_from` -- exactly the bytes the receiver needs (one ncclSend + one ncclRecv in a group), instead of the
    N-way all-gather whose receive volume grows with the world size.  Returns `out` (and the list of works when async_op=True;
    `wait_all(works)` orders the current stream after them)."""
    if out is None:
        out = torch.empty_like(offer)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out.copy_(offer)
        return (out, []) if async_op else out
    ops = [dist.P2POp(dist.isend, offer.contiguous(), send_to), dist.P2POp(dist.irecv, out, recv_from)]
    works = dist.batch_isend_irecv(ops)
    if async_op:
        return out, works
    wait_all(works)
    return out


def wait_all(works):
    for w in works or []:
        w.wait()


def exchange_requested(clouds, pairs, rank=None, world=None, device=None):
    """General form of the loop-closure exchange: `clouds` maps this robot's keyframe index -> tensor [N, 3];
    `pairs` is the GLOBAL candidate list (src_robot, src_kf, tgt_robot, tgt_kf).  Every rank sends each peer only the target
    clouds that peer's pairs ask for (one packed buffer per peer, grouped ncclSend/ncclRecv) and gets back a dict
    (tgt_robot, tgt_kf) -> tensor for the pairs it refines.  All clouds must have the same shape."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    want = {}     # peer -> sorted list of (robot, kf) I need from it
    give = {}     # peer -> sorted list of kf it needs from me
    for p in pairs:
        a, b = owner_of_pair(p, world), p[2] % world
        if a == b:
            continue
        if a == rank:
            want.setdefault(b, set()).add((p[2], p[3]))
        if b == rank:
            give.setdefault(a, set()).add(p[3])
    want = {k: sorted(v) for k, v in want.items()}
    give = {k: sorted(v) for k, v in give.items()}
    any_cloud = next(iter(clouds.values()))
    ops, recv_bufs, keep = [], {}, []
    for peer, kfs in sorted(give.items()):
        buf = torch.stack([clouds[k] for k in kfs]).contiguous()
        keep.append(buf)
        ops.append(dist.P2POp(dist.isend, buf, peer))
    for peer, keys in sorted(want.items()):
        buf = torch.empty((len(keys),) + tuple(any_cloud.shape), dtype=any_cloud.dtype, device=any_cloud.device if device is None else device)
        recv_bufs[peer] = buf
        ops.append(dist.P2POp(dist.irecv, buf, peer))
    if ops:
        wait_all(dist.batch_isend_irecv(ops))
    got = {}
    for peer, keys in want.items():
        for j, key in enumerate(keys):
            got[key] = recv_bufs[peer][j]
    for p in pairs:   # targets that live on this rank need no exchange
        if owner_of_pair(p, world) == rank and p[2] % world == rank:
            got[(p[2], p[3])] = clouds[p[3]]
    return got


def shard_batch(n_items, rank, world):
    """Contiguous split of a batch of independent units (RING scans): [lo, hi) of this rank."""
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


def allgather_descriptors(local_desc, n_total):
    """RING multi-GPU step: every rank has built the descriptors of its shard of the scan batch (`shard_batch`);
    one all-gather makes the whole database [n_total, ...] available on every rank, which then correlates its own queries
    against all of it.  `local_desc` is a real tensor [n_local, ...] (complex descriptors as view_as_real); shards may differ
    by one scan, so they are padded to the largest."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_desc
    world = dist.get_world_size()
    nmax = (n_total + world - 1) // world
    pad = torch.zeros((nmax,) + tuple(local_desc.shape[1:]), dtype=local_desc.dtype, device=local_desc.device)
    pad[: local_desc.shape[0]] = local_desc
    out = torch.empty((world,) + tuple(pad.shape), dtype=pad.dtype, device=pad.device)
    if pad.is_cuda:
        dist.all_gather_into_tensor(out, pad)
    else:
        dist.all_gather(list(out.unbind(0)), pad)
    parts = []
    for r in range(world):
        lo, hi = shard_batch(n_total, r, world)
        parts.append(out[r, : hi - lo])
    return torch.cat(parts)


def gather_results(local_T, local_ids, n_total):
    """local_T: [n_local,4,4] float32, local_ids: global pair indices.  Returns [n_total,4,4] on every rank."""
    T = torch.zeros((n_total, 4, 4), dtype=torch.float32, device=local_T.device)
    if len(local_ids):
        T[torch.as_tensor(local_ids, device=local_T.device, dtype=torch.long)] = local_T
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(T, op=dist.ReduceOp.SUM)  # disjoint supports: a sum is a gather
    return T


def ring_candidate_pairs(world_size, per_rank):
    """The benchmark's loop-closure storm (BASELINE config 4): robot r's keyframe i is matched against keyframe i of
    robot (r+1) % world.  Returns the global pair list."""
    return [(r, i, (r + 1) % world_size, i) for r in range(world_size) for i in range(per_rank)]


def bind_to_gpu_numa(local_rank):
    """Pin the calling process to the CPU cores NVML reports as local to GPU `local_rank` (nvmlDeviceSetCpuAffinity).  One process
    per GPU on a two-socket box otherwise lands on either socket; the LM driver synchronises with the device a few dozen times per
    batch and reads 28 doubles per problem out of pinned memory each time, so a remote socket costs latency on every one of them.
    Returns the core list, or None when NVML is not available (nothing is changed then)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return None


class TiledElevationMap:
    """Tile-sharded global elevation grid (BASELINE config 5): rank r owns a block of STORAGE rows of one L x L grid and
    fuses only the points that fall into it (`set_owned_rows`); every rank scans all points, cells are disjoint across tiles, so the
    union of the tiles equals the single-GPU map bit for bit and the fuse needs no reduction.  Only neighbourhood operations need
    data of other tiles: before Map_feature the `halo` (= 2, the 5 x 5 window) storage rows on either side of the block are fetched
    from the neighbouring ranks (one small all_gather of the border rows).  Raytracing walks rays across the whole grid and is not
    tiled.

    `emap` is an initialised map object with the product interface (mr_slam_b200.elevation.ElevationMap; the CPU oracle twin in the
    gloo tests); `device` is where the exchange tensors live ("cuda" for NCCL, "cpu" for gloo); `sync` is called between library-
    stream copies and collectives (the two use different streams)."""

    def __init__(self, emap, length, rank=None, world=None, device="cuda", sync=None, halo=2):
        self.m, self.L, self.halo, self.device = emap, int(length), int(halo), device
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.bounds = [(r * self.L) // self.world for r in range(self.world + 1)]
        self.row_lo, self.row_hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        if self.world > 1 and min(b - a for a, b in zip(self.bounds[:-1], self.bounds[1:])) < self.halo:
            raise ValueError("tiles must be at least `halo` rows high")
        self.stream = None
        if sync is None and str(device).startswith("cuda"):
            # rows_to_device / rows_from_device enqueue D2D copies on the LIBRARY stream.  The torch ops and the collectives of the
            # exchange are issued with that same stream as torch's current stream, so everything is ordered by the stream itself --
            # no host synchronisation (the first version synchronised the device three times per exchange: 0.6 ms at 4000^2).
            from ._lib import lib
            self.stream = torch.cuda.ExternalStream(lib().mrslam_stream())
        self.sync = sync if sync is not None else (lambda: None)
        emap.set_owned_rows(self.row_lo, self.row_hi)

    def _spans(self, row, n):
        """storage rows wrap around (ring buffer): [row, row+n) as at most two contiguous spans (start, count, offset)"""
        row %= self.L
        first = min(n, self.L - row)
        return [(row, first, 0)] + ([(0, n - first, first)] if n > first else [])

    def _rows(self, plane, row, n, dtype):
        t = torch.empty((n, self.L), dtype=dtype, device=self.device)
        for start, cnt, off in self._spans(row, n):
            self.m.rows_to_device(plane, start, cnt, t[off:off + cnt])
        return t

    def _put(self, plane, row, t):
        t = t.contiguous()
        for start, cnt, off in self._spans(row, t.shape[0]):
            self.m.rows_from_device(plane, start, cnt, t[off:off + cnt])

    def _on_stream(self):
        import contextlib
        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def exchange_halo(self, plane="elevation"):
        """fetch the `halo` storage rows above and below the own block from the neighbouring tiles into this replica"""
        if self.world == 1:
            return
        with self._on_stream():
            self._exchange_halo(plane)

    def _exchange_halo(self, plane):
        h = self.halo
        dtype = torch.int32 if plane in ("R", "G", "B") else torch.float32
        mine = torch.cat([self._rows(plane, self.row_lo, h, dtype), self._rows(plane, self.row_hi - h, h, dtype)])   # top, bottom
        self.sync()
        if mine.is_cuda:
            allt = torch.empty((self.world,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
            dist.all_gather_into_tensor(allt, mine)
            allb = list(allt.unbind(0))
        else:
            allb = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allb, mine)
        self.sync()
        prev_r, next_r = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        self._put(plane, self.row_lo - h, allb[prev_r][h:])      # the rows just above my block are my predecessor's bottom rows
        self._put(plane, self.row_hi, allb[next_r][:h])          # the rows just below are my successor's top rows
        self.sync()

    def map_feature_owned(self):
        """Map_feature for the owned rows: dict of [rows, L] arrays (numpy) for elevation / var / rough / slope / traver / ..."""
        self.exchange_halo("elevation")
        out = self.m.Map_feature(self.L) if hasattr(self.m, "Map_feature") else self.m.map_feature()
        return {k: np.asarray(v).reshape(self.L, self.L)[self.row_lo:self.row_hi] for k, v in out.items()}

    def gather_full(self, planes=("elevation", "variance")):
        """assemble whole planes on every rank from the owned blocks (all_gather of padded blocks); returns {plane: tensor [L*L]}"""
        with self._on_stream():
            res = self._gather_full(planes)
        if self.stream is not None:
            self.stream.synchronize()     # the caller reads the result on its own stream / the host
        return res

    def _gather_full(self, planes):
        res = {}
        hmax = max(b - a for a, b in zip(self.bounds[:-1], self.bounds[1:]))
        for plane in planes:
            dtype = torch.int32 if plane in ("R", "G", "B") else torch.float32
            blk = torch.zeros((hmax, self.L), dtype=dtype, device=self.device)
            n = self.row_hi - self.row_lo
            if n:
                self.m.rows_to_device(plane, self.row_lo, n, blk[:n])
            self.sync()
            if self.world > 1:
                allb = [torch.empty_like(blk) for _ in range(self.world)]
                dist.all_gather(allb, blk)
            else:
                allb = [blk]
            self.sync()
            res[plane] = torch.cat([allb[r][: self.bounds[r + 1] - self.bounds[r]] for r in range(self.world)]).reshape(-1)
        return res

