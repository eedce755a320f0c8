"""The path's collectives: the batch is sharded by rank with no data-path exchange;
the only cross-rank data are the int64 PCK/AUC/EPE counters and the four f64 loss sums.

Mirrors the reference's helpers (all of which are unused call sites there):
  train/spawn_dist.py:68-80          all_reduce(values)   list -> f32 tensor -> barrier -> SUM
  train/distributed_utils.py:65-76   reduce_value(value, average=True)
Backend: NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests.  No barrier is issued
before the all-reduce (the reference's dist.barrier() only adds latency).
"""
import torch
import torch.distributed as dist


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_bounds(n, rank, world):
    """Contiguous batch shard [lo, hi) of rank `rank`."""
    return (n * rank) // world, (n * (rank + 1)) // world


def all_reduce(values, device=None):
    """spawn_dist.py:68-80: list of python numbers -> summed f32 tensor (on every rank)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
    t = torch.tensor(values, dtype=torch.float32, device=device)
    if _active():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def reduce_value(value, average=True):
    """distributed_utils.py:65-76: mean (or sum) of a tensor over the ranks (DDP loss logging)."""
    if not _active():
        return value
    with torch.no_grad():
        value = value.clone()
        dist.all_reduce(value)
        if average:
            value /= dist.get_world_size()
    return value


def all_reduce_loss_sums(sums):
    """Global-batch semantics of the balanced loss: SUM the f64 (S_pos, S_neg, N_pos, numel) of every
    shard, then finalise — equals the single-process reference on the concatenated batch."""
    if _active():
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


def all_reduce_counters(counters):
    """int64 SUM of the metric counters: bit-exact for any sharding."""
    if _active():
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


class PeerExchange:
    """Mailboxes of the in-kernel exchange (include/lhn.h lhn_exchange): every rank owns LHN_XCH_MAILBOX_BYTES of
    zeroed device memory that is mapped into every peer of the node — torch symmetric memory (CUDA VMM handles
    exchanged over the process group's store), or cudaIpc handles as the fallback.  The kernels of
    lhn_fused_render_loss_decode_xch / lhn_decode_heatmap_pck_xch store their step's block straight into the peers'
    mailboxes over NVLink and add what arrives, so the per-step all-reduce of SURVEY §8e needs no NCCL call and no SM
    left free for one.  `next_seq()` numbers the steps; every rank must issue the same sequence of exchanging launches.
    """

    def __init__(self, device, group=None):
        from . import _lib as L
        self.device = torch.device(device)
        self.world = dist.get_world_size(group) if _active() else 1
        self.rank = dist.get_rank(group) if _active() else 0
        if self.world > L.XCH_MAX_RANKS:
            raise L.LhnError(f"in-kernel exchange supports up to {L.XCH_MAX_RANKS} ranks of one node")
        self.seq = 0
        self.how = "local"
        self._keep = []
        self._blocks, self._pending = None, None
        if self.world == 1:
            self.mailbox = torch.zeros(L.XCH_MAILBOX_BYTES, dtype=torch.uint8, device=self.device)
            self.ptrs = [self.mailbox.data_ptr()]
        else:
            try:
                self.ptrs = self._symmetric(L.XCH_MAILBOX_BYTES, group)
                self.how = "torch symmetric memory (CUDA VMM, peer-mapped over NVLink)"
            except Exception as e:                                   # pragma: no cover - depends on the driver stack
                self._symm_error = f"{type(e).__name__}: {e}"
                self.ptrs = self._ipc(L.XCH_MAILBOX_BYTES, group)
                self.how = "cudaIpc memory handles (peer-mapped over NVLink)"
            torch.cuda.synchronize(self.device)
            dist.barrier(group)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _symmetric(self, nbytes, group):
        import torch.distributed._symmetric_memory as symm_mem
        g = group if group is not None else dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(g.group_name)
        except Exception:
            pass
        self.mailbox = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.mailbox.zero_()
        torch.cuda.synchronize(self.device)
        hdl = symm_mem.rendezvous(self.mailbox, g)
        self._keep.append(hdl)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.mailbox.data_ptr():
            raise RuntimeError("unexpected symmetric-memory handle layout")
        return ptrs

    def _ipc(self, nbytes, group):
        self.mailbox = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        torch.cuda.synchronize(self.device)
        handle = self.mailbox.untyped_storage()._share_cuda_()
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(self.mailbox.data_ptr())
                continue
            st = torch.UntypedStorage._new_shared_cuda(*h)
            t = torch.empty(0, dtype=torch.uint8, device=st.device).set_(st)
            self._keep.append(t)
            ptrs.append(t.data_ptr())
        return ptrs

    @classmethod
    def local_group(cls, n, device):
        """n mailboxes on ONE device wired to each other (protocol tests: n 'ranks' on n streams of a single GPU)."""
        from . import _lib as L
        boxes = [torch.zeros(L.XCH_MAILBOX_BYTES, dtype=torch.uint8, device=device) for _ in range(n)]
        out = []
        for r in range(n):
            x = cls.__new__(cls)
            x.device, x.world, x.rank, x.seq, x.how = torch.device(device), n, r, 0, "local test group"
            x._blocks, x._pending = None, None
            x.mailbox, x.ptrs, x._keep = boxes[r], [b.data_ptr() for b in boxes], boxes
            x.status = torch.zeros(1, dtype=torch.int32, device=device)
            out.append(x)
        return out

    def next_seq(self):
        self.seq += 1
        return self.seq

    # ---- per-step counter blocks of the pipelined exchange (lhn_decode_heatmap_pck_xch) ---------------------------
    def step_blocks(self, n):
        """[LHN_XCH_SLOTS, n] int64 zeros: the launch of step s accumulates into row s % LHN_XCH_SLOTS; the next
        exchanging launch sends it to the peers, the one after adds all ranks' rows into the totals and zeroes it
        (flush() does both for what is still in flight)."""
        from . import _lib as L
        if self._blocks is None or self._blk_n != n:
            if self._pending:
                raise L.LhnError("flush() the exchange before changing the block size")
            # rows padded to an even word count: the kernel moves them with 16-byte-granular bulk copies
            self._blocks = torch.zeros((L.XCH_SLOTS, (n + 1) & ~1), dtype=torch.int64, device=self.device)
            self._blk_ptrs = [self._blocks[i].data_ptr() for i in range(L.XCH_SLOTS)]
            self._blk_n = n
        return self._blocks

    def begin_step(self, n, totals, x):
        """Number a new exchanging launch and fill its lhn_exchange `x`: this step's sequence number, the previous
        step's block (this launch publishes it) and the block of the step before (this launch adds it into the totals).
        Returns this step's block pointer.  (Runs once per launch on the host: kept to a few attribute stores.)"""
        if self._blocks is None or self._blk_n != n:
            self.step_blocks(n)
        seq = self.seq = self.seq + 1
        cur = self._blk_ptrs[seq & 3]
        p = self._pending
        x.seq = seq
        if p:
            x.prev_block, x.prev_seq = p[-1][1], p[-1][0]
            if len(p) > 1:
                x.prev2_block, x.prev2_seq = p[0][1], p[0][0]
                self._pending = [p[1], (seq, cur)]
            else:
                x.prev2_block, x.prev2_seq = None, 0
                self._pending = [p[0], (seq, cur)]
        else:
            x.prev_block, x.prev_seq, x.prev2_block, x.prev2_seq = None, 0, None, 0
            self._pending = [(seq, cur)]
        self._pending_meta = (n, totals)
        return cur

    def flush(self, timeout_ms=2000):
        """Complete the (up to two) steps still in flight (lhn_exchange_flush) — after it `totals` holds every step of
        every rank, on every rank."""
        from . import _lib as L
        if not self._pending:
            return
        n, totals = self._pending_meta
        p = self._pending
        x = self.struct(timeout_ms)
        x.seq = p[-1][0]
        x.prev_block, x.prev_seq = p[-1][1], p[-1][0]
        x.prev2_block, x.prev2_seq = (p[-2][1], p[-2][0]) if len(p) >= 2 else (None, 0)
        with L.on_device(self.device):
            L.check(L.lib().lhn_exchange_flush(x, n, L.ptr(totals), L.stream(self.device)), "lhn_exchange_flush")
        self._pending = None

    def struct(self, timeout_ms=2000):
        """A fresh lhn_exchange for one bound launcher (its seq is set right before every launch)."""
        from . import _lib as L
        x = L.Exchange()
        for r, p in enumerate(self.ptrs):
            x.mailbox[r] = p
        x.world, x.rank, x.seq, x.timeout_ms = self.world, self.rank, 0, int(timeout_ms)
        x.status = self.status.data_ptr()
        return x

    def bytes_per_step(self, payload_bytes):
        """NVLink bytes this rank sends per exchanging launch: the block + a 4-byte flag to each peer."""
        return (self.world - 1) * (int(payload_bytes) + 4)
