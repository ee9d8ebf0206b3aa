"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

How the path shards:
  * by SCENE (default; the reference's only strategy is DDP, train.py:90-131): every rank owns whole samples,
    no data-path collective, weak scaling.  `scene_partition` gives each rank its sample indices.
  * by FRAME, features replicated (partitioning A, the north star's wording): rank r's backbone produces frames
    [r*T/N, (r+1)*T/N); `all_gather_features` (one NCCL all-gather per FPN level per forward) gives every rank the whole
    pyramid and the decoder runs unsharded (replicated, or scene-/query-parallel on top).
  * by FRAME, features local (partitioning B): rank r owns frames [r*T/N, (r+1)*T/N) of every scene -- features never cross
    NVLink; each rank gathers its own frames for all queries and ONE all-gather per layer assembles the
    sampled features `[T, B, Q, G, P, C]` (frame-major staging so every rank's chunk is contiguous).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (no-op for world size 1)."""
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        kw = {'device_id': device} if (backend == 'nccl' and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device='cpu'):
    """Device-timed durations are reported as the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def scene_partition(num_scenes, rank, world):
    """Contiguous block of scene indices for this rank (sizes differ by at most one)."""
    base, rem = divmod(num_scenes, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def frame_partition(num_frames, rank, world):
    """Frames [t0, t1) owned by this rank; num_frames must divide evenly so the all-gather chunks are equal."""
    if num_frames % world:
        raise ValueError('num_frames (%d) must be divisible by the world size (%d)' % (num_frames, world))
    per = num_frames // world
    return rank * per, (rank + 1) * per


def all_gather_frames(local, world=None):
    """local [T_local, ...] on every rank -> [T, ...] (frame-major, rank order == frame order)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = world or dist.get_world_size()
    out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def all_gather_features(mlvl_local, group=None):
    """Partitioning A of SURVEY.md 8(e) (the north star's "NCCL all-gather of per-camera features"): every rank's
    backbone produced the FPN levels of ITS frames only, `[B, Tl*N, C, H, W]` per level (frame-major camera axis as in
    the reference, models/sparsebev.py:124-131; plain NCHW or channels-last memory, e.g. backbone.extract_img_feat);
    returns the levels of all T = world*Tl frames on every rank, `[B, T*N, C, H, W]`, in the SAME memory format --
    ready for SparseBEVTransformerDecoder.prepare_feats (channels-last stays zero-copy).  One collective per level per
    forward (92 MB per frame at r50, 524 MB at vov99); for B == 1 the gathered buffer already is the result (rank order
    == frame order), for B > 1 one re-layout copy moves the batch axis back in front."""
    if not (dist.is_available() and dist.is_initialized()):
        return list(mlvl_local)
    world = dist.get_world_size(group)
    if world == 1:
        return list(mlvl_local)
    out = []
    for feat in mlvl_local:
        nhwc = feat.permute(0, 1, 3, 4, 2)
        cl = nhwc.is_contiguous() and not feat.is_contiguous()
        x = nhwc if cl else feat.contiguous()                                # [B, Tl*N, ...] in its memory order
        B = x.shape[0]
        buf = x.new_empty((world,) + tuple(x.shape))                          # [W, B, Tl*N, ...]
        dist.all_gather_into_tensor(buf.view((world * B,) + tuple(x.shape[1:])), x, group=group)
        if B == 1:
            full = buf.view((1, world * x.shape[1]) + tuple(x.shape[2:]))
        else:
            full = buf.transpose(0, 1).reshape((B, world * x.shape[1]) + tuple(x.shape[2:]))
        out.append(full.permute(0, 1, 4, 2, 3) if cl else full)
    return out


def merge_frame_chunks(chunks):
    """[W, B, Q, G, Tl*P, C] (rank-major, as all_gather_into_tensor returns it) -> [B, Q, G, W*Tl*P, C].

    Rank r owns frames [r*Tl, (r+1)*Tl), so point index t*P + p of the full tensor is r*(Tl*P) + (t - r*Tl)*P + p."""
    W, B, Q, G, TP, C = chunks.shape
    return chunks.permute(1, 2, 3, 0, 4, 5).reshape(B, Q, G, W * TP, C)


class FrameShard:
    """Frame-sharded decoder state of one rank (partitioning B).

    Every rank keeps the feature maps of the frames [t0, t1) its backbone produced, runs the (tiny) query-side
    stages redundantly, samples ITS frames for all queries, and the sampled rows are exchanged once per layer:

      exchange='p2p'  (default on GPUs): two symmetric-memory buffers [B,Q,G,T*P,C] per rank, mapped into every peer.
                      The gather kernel stores each 256 B row into all ranks' buffers over NVLink
                      (sbev_sampling4d_scatter_fwd: compute and all-gather are one kernel), then ONE cross-GPU barrier
                      on the stream; buffers alternate between layers so a rank never overwrites rows a slower
                      peer is still mixing (a rank is at most one barrier ahead of any peer).
      exchange='nccl' local [B,Q,G,Tl*P,C] -> all_gather_into_tensor -> merge_frame_chunks (one re-layout copy).
                      Also the path the gloo CPU tests drive.
    """

    def __init__(self, num_frames, rank=None, world=None, group=None, exchange='p2p'):
        if rank is None or world is None:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError('FrameShard needs an initialised process group (or explicit rank/world)')
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        if exchange not in ('p2p', 'nccl'):
            raise ValueError('exchange must be "p2p" or "nccl"')
        self.rank, self.world, self.group, self.exchange = rank, world, group, exchange
        self.num_frames = num_frames
        self.window = frame_partition(num_frames, rank, world)
        self._bufs, self._turn = {}, 0

    # ---- nccl / gloo exchange
    def all_gather(self, local):
        """local [B,Q,G,Tl*P,C] -> [B,Q,G,T*P,C]."""
        if self.world == 1:
            return local
        chunks = local.new_empty((self.world * local.shape[0],) + tuple(local.shape[1:]))     # rank-major concatenation
        dist.all_gather_into_tensor(chunks, local.contiguous(), group=self.group)
        return merge_frame_chunks(chunks.view((self.world,) + tuple(local.shape)))

    # ---- peer-store exchange
    def peer_buffers(self, shape, device):
        """-> (this rank's full-size buffer for the current layer, [device address of that buffer on every rank])."""
        import torch.distributed._symmetric_memory as symm_mem
        key = (tuple(shape), str(device))
        if key not in self._bufs:
            pair = []
            for _ in range(2):
                buf = symm_mem.empty(*shape, dtype=torch.float32, device=device)
                hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
                pair.append((buf, hdl, [int(p) for p in hdl.buffer_ptrs]))
            self._bufs[key] = pair
        buf, hdl, ptrs = self._bufs[key][self._turn]
        self._cur = hdl
        self._turn ^= 1
        return buf, ptrs

    def peer_barrier(self):
        """Stream-ordered barrier over all ranks: returns (on the stream) once every rank's rows have landed."""
        self._cur.barrier(channel=0)


def query_partition(num_query, rank, world):
    """-> (q_per_rank, q0, q1): rank r owns queries [r*q_per_rank, min((r+1)*q_per_rank, Q)), q_per_rank = ceil(Q / world)
    (the last ranks may own fewer -- or none, when Q < world)."""
    qpr = -(-num_query // world)
    q0 = min(rank * qpr, num_query)
    return qpr, q0, min(q0 + qpr, num_query)


class QueryShard:
    """Query- AND frame-sharded decoder state of one rank: ONE scene (B = 1) across `world` GPUs, strong scaling
    (frames local as in partitioning B, every query-side stage sharded over queries).

    Rank r keeps the feature maps of frames [t0, t1) (what its backbone produced; the pyramid never crosses NVLink) and
    OWNS queries [q0, q1).  Per decoder layer:
      * position encoder + attention in-projection run for all Q rows on every rank (K / V of every query are needed);
      * attention core, out-projection, sampling heads, sample points: own queries only;
      * exchange 1 (sbev_peer_exchange): every rank's sample points + scale weights -> all ranks (0.4 MB at Q = 900);
      * gather: own frames, ALL queries; each 256 B row is stored straight into the buffer of the rank that owns the
        query (sbev_sampling4d_owner_fwd: the all-to-all of sampled rows is the gather's store), then a barrier;
      * parameter GEMM, mixing, out-projection, FFN, cls / reg branches, box refinement: own queries only;
      * exchange 2: refined boxes, query features, class scores of the own queries -> all ranks (1 MB at Q = 900).
    Every exchange is one kernel of ours that stores into the peers' symmetric memory over NVLink and ends in an
    all-ranks barrier on flag words in that memory; no NCCL on the data path.  All buffers are written at most once
    between two barriers and read only after the barrier that follows the write, so nothing is double-buffered except
    the layer outputs (a layer reads the previous layer's output buffers while it fills the other set).
    """

    FLAG_WORDS = 64            # 256 B at the start of the arena: SBEV_MAX_PEERS flag words + padding

    def __init__(self, num_frames, rank=None, world=None, group=None, emulate=False):
        """emulate=True (development / profiling on ONE GPU): this process plays rank `rank` of `world` without peers -- the
        arena is ordinary device memory, every "peer" buffer aliases it and the exchanges degenerate to the local part of
        the kernel, so the rank-local kernel sequence of an N-GPU run can be profiled (ncu) on a single GPU.  Results are
        NOT the sharded decoder's (rows of the other ranks never arrive)."""
        self.emulate = bool(emulate)
        if rank is None or world is None:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError('QueryShard needs an initialised process group (or explicit rank/world)')
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world > 8:
            raise ValueError('QueryShard supports at most 8 ranks (one NVSwitch node)')
        self.rank, self.world, self.group = rank, world, group
        self.num_frames = num_frames
        self.window = frame_partition(num_frames, rank, world)
        self._arenas = {}
        self._parity = 0
        self.exchanges = 0         # exchanges enqueued (host-side count; every rank must issue the same sequence)
        self.bytes_sent = 0        # payload bytes this rank stored into peers through sbev_peer_exchange (host-side count)

    def partition(self, num_query):
        return query_partition(num_query, self.rank, self.world)

    @staticmethod
    def layout(fields):
        """fields [(name, shape)] -> ({name: (offset_in_floats, shape)}, total_floats); every field starts on a 256-byte boundary."""
        off, table = QueryShard.FLAG_WORDS, {}
        for name, shape in fields:
            n = 1
            for s in shape:
                n *= int(s)
            table[name] = (off, tuple(int(s) for s in shape))
            off += (n + 63) // 64 * 64
        return table, off

    def arena(self, fields, device):
        """One symmetric-memory allocation per distinct field list: local tensor views + every rank's base address."""
        key = (tuple((n, tuple(s)) for n, s in fields), str(device))
        ar = self._arenas.get(key)
        if ar is None and self.emulate:
            table, total = self.layout(fields)
            buf = torch.zeros(total, dtype=torch.float32, device=device)
            views = {n: buf[o:o + _numel(s)].view(s) for n, (o, s) in table.items()}
            ar = dict(buf=buf, hdl=None, ptrs=[buf.data_ptr()] * self.world, table=table, views=views,
                      ctl=torch.zeros(4, dtype=torch.int32, device=device), device=device)
            self._arenas[key] = ar
        if ar is None:
            import torch.distributed._symmetric_memory as symm_mem
            table, total = self.layout(fields)
            buf = symm_mem.empty(total, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            buf.zero_()
            ctl = torch.zeros(4, dtype=torch.int32, device=device)
            torch.cuda.synchronize(device)
            dist.barrier(group=self.group)          # every rank's flag words are zero before anyone signals
            views = {n: buf[o:o + _numel(s)].view(s) for n, (o, s) in table.items()}
            ar = dict(buf=buf, hdl=hdl, ptrs=[int(p) for p in hdl.buffer_ptrs], table=table, views=views, ctl=ctl, device=device)
            self._arenas[key] = ar
        return ar

    def peer_ptrs(self, ar, name):
        off = ar['table'][name][0]
        return [p + 4 * off for p in ar['ptrs']]

    def exchange(self, ar, rows):
        """rows = [(field, r0, r1)]: rows [r0, r1) of every field (leading dimension) go to all peers; then the barrier."""
        from . import ops
        segs = []
        for name, r0, r1 in rows:
            off, shape = ar['table'][name]
            per = _numel(shape[1:])
            start, nbytes = 4 * (off + r0 * per), 4 * max(0, r1 - r0) * per
            segs.append((ar['ptrs'][self.rank] + start, [p + start for p in ar['ptrs']], nbytes))
            self.bytes_sent += nbytes * (self.world - 1)
        self.exchanges += 1
        if self.emulate:             # no peers: the kernel's local part only (arrival counter + epoch), nothing to copy or wait for
            ops.peer_exchange([], 1, 0, ar['ptrs'][:1], ar['ctl'].data_ptr(), ar['device'])
            return
        ops.peer_exchange(segs, self.world, self.rank, ar['ptrs'], ar['ctl'].data_ptr(), ar['device'])

    def flip(self):
        self._parity ^= 1
        return self._parity

    def status(self, ar):
        """0 = healthy; 1 = an exchange gave up waiting for a peer (synchronises the stream)."""
        return int(ar['ctl'][2].item())


def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n
