"""``mc_predict`` - the fused replacement of the reference's S-pass loop, plus sample sharding.

Reference: ``FullAnalysis._get_output`` (Software_Artifact/software/train/results_analyzer.py:236-270).
Multi-GPU (new; the reference is single-device): rank r of G owns the contiguous global samples
``shard_samples(S, G, r)``; the deterministic prefix is replicated; ONE all-reduce(sum) of the flat
per-exit statistics buffer ([E,B,C] sum of probs, [E,B,C] sum of logits, [E,B] sum of p log p) over
NCCL precedes the finaliser.  Because Philox counters carry the GLOBAL sample index the reduced sums
do not depend on G.
"""
import torch

from .engine import MCResult  # noqa: F401


def shard_samples(S, world_size, rank):
    """(first global sample, number of local samples) of `rank`: contiguous, sizes differ by <= 1."""
    if S < 0 or world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad shard request S=%d world=%d rank=%d" % (S, world_size, rank))
    base, extra = divmod(S, world_size)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_batch(B, world_size, rank):
    """Alternative partitioning (independent images): contiguous slice of the batch."""
    return shard_samples(B, world_size, rank)


class PeerReduce:
    """The path's single collective - the sum of the flat statistics buffer over the ranks - as ONE kernel over NVLink
    peer memory (csrc/kernels_comm.cu, `bnn_peer_allreduce`): every rank copies its sums into a slot of symmetric memory
    (torch.distributed._symmetric_memory maps it into all peers), signals the peers, adds the W slots in RANK order - so
    the totals are bit-identical on every rank - and waits until all peers have read its slot.  Stream-ordered, CUDA-graph
    capturable, no host synchronisation.  One workspace per payload size; creating it is a collective (first call)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self._ws = {}
        self.disabled = None                     # a string: why NCCL is used instead

    def _workspace(self, flat):
        import ctypes
        import torch.distributed as dist
        key = (flat.numel(), flat.device.index)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        ok, why = 1, ""
        try:
            import torch.distributed._symmetric_memory as symm
            n_pad = (flat.numel() + 3) // 4 * 4
            t = symm.empty(n_pad + 16, dtype=torch.float32, device=flat.device)     # slot | ready[8] | done[8]
            t.zero_()
            hdl = symm.rendezvous(t, self.group)
            bases = (ctypes.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            state = torch.zeros(8, dtype=torch.int32, device=flat.device)
            ws = (t, hdl, bases, state, n_pad * 4)
        except Exception as e:                     # noqa: BLE001 - no P2P mapping on this box / torch build
            ok, why = 0, "%s: %s" % (type(e).__name__, e)
        # all ranks take the same decision, and nobody signals before every rank's flags are zero
        flag = torch.tensor([ok], dtype=torch.int32, device=flat.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            import warnings
            self.disabled = why or "a peer could not map the symmetric workspace"
            warnings.warn("peer-memory all-reduce unavailable (%s): using NCCL" % self.disabled)
            return None
        self._ws[key] = ws
        return ws

    def __call__(self, flat):
        import ctypes
        import torch.distributed as dist
        from . import _lib
        if self.disabled is None and flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous():
            ws = self._workspace(flat)
            if ws is not None:
                _, _, bases, state, flags_off = ws
                lib = _lib.load()
                _lib.check(lib.bnn_peer_allreduce(ctypes.c_void_p(flat.data_ptr()), flat.numel(), bases, self.world, self.rank,
                                                  flags_off, ctypes.c_void_p(state.data_ptr()),
                                                  ctypes.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)))
                return flat
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat

    def check(self):
        """Raise if a peer ever failed to answer (synchronises the stream)."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        for (_, dev), (_, _, _, state, _) in self._ws.items():
            _lib.check(lib.bnn_peer_allreduce_status(ctypes.c_void_p(state.data_ptr()),
                                                     ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))


_PEER = {}           # process group -> PeerReduce


def allreduce_sums_nccl(flat, group=None):
    """The same sum through NCCL (`BNN_PEER_REDUCE=0`, CPU / gloo groups, and the baseline bench.py times beside it)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def allreduce_sums(flat, group=None):
    """The single collective of the path: sum the flat statistics buffer over ranks (in place, on the current stream)."""
    import os
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return flat
    if flat.is_cuda and os.environ.get("BNN_PEER_REDUCE", "1") != "0" and dist.get_world_size(group) <= 8:
        g = group if group is not None else dist.group.WORLD
        red = _PEER.get(g)
        if red is None:
            red = _PEER[g] = PeerReduce(g)
        return red(flat)
    return allreduce_sums_nccl(flat, group)


def _generic_engine(model, x, dtype, engine_kwargs):
    """Plan for an arbitrary nn.Module (lowering.lower_module), cached per (model, input shape, dtype) in `_plans`."""
    from . import engine, lowering
    flat = x.dim() == 2
    shape = (int(x.shape[1]), 1, 1) if flat else tuple(int(v) for v in x.shape[1:])
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("bayesnn_fpga_b200 runs on a CUDA (sm_100) device only; there is no CPU fallback - call "
                           "model.cuda() first")
    from . import _plans
    cache = _plans.plans_for(model)          # kept outside the (foreign) module; re-validated against its parameters
    key = ("generic", shape, dtype, tuple(sorted((engine_kwargs or {}).items())))
    if key not in cache:
        graph, _ = lowering.lower_module(model, shape)
        cache[key] = engine.Engine(graph, dtype=dtype or "fp16", device=dev, **(engine_kwargs or {}))
    return cache[key], shape


def mc_predict(model, x, S, seed=0x5EED, dtype=None, want_logits=False, distributed=False, group=None,
               engine_kwargs=None):
    """S stochastic passes of `model` over the batch `x` ([B, C, H, W] float32, host or device).

    Returns an :class:`MCResult` whose tensors live in the engine's cached device buffers (valid
    until the next call on the same model/batch shape; ``.clone()`` to keep).  With
    ``distributed=True`` (inside an initialised ``torch.distributed`` process group) the S samples
    are sharded over the ranks and every rank gets the full statistics.
    """
    if S <= 0:
        raise ValueError("S must be positive, got %d" % S)
    if hasattr(model, "bnn_engine"):
        eng = model.bnn_engine(dtype, **(engine_kwargs or {}))
    else:
        # any other nn.Module (e.g. the output of nn2bnn._convert_model, or a user network holding MCDropout /
        # Masksembles modules): traced and lowered to the same plan; exits = the module's outputs in return order
        eng, shape = _generic_engine(model, x, dtype, engine_kwargs)
        x = x.reshape(x.shape[0], *shape)
    sample0, s_local, reduce_fn = 0, S, None
    if distributed:
        import torch.distributed as dist
        ws, rk = dist.get_world_size(group), dist.get_rank(group)
        sample0, s_local = shard_samples(S, ws, rk)
        reduce_fn = lambda flat: allreduce_sums(flat, group)
    return eng.run(x, s_local, seed=seed, sample0=sample0, S_total=S, want_logits=want_logits,
                   reduce_fn=reduce_fn)


def masksembles_batched_predict(model, x, dtype=None):
    """The Keras converter's Masksembles inference (``MasksemblesModel.call(training=False)``,
    Hardware_Artifact/converter/keras/Masksembles.py:216-239) for this package's multi-exit models: every input runs
    through ALL n masks in one call and the n predictions are averaged.  The reference tiles the input n times along the
    batch and lets every layer split it into n groups (:177-181, :220-223); here that is S = n samples of the fused plan
    with sample g using mask row g - the layers in front of the first Masksembles layer run ONCE per image instead of n
    times, nothing else changes.  The modules' eval-mode rotation counters are neither used nor advanced.

    -> dict: ``per_exit`` [E, B, C] = mean over the n masks of each exit's soft-max output (:229-232),
             ``prediction`` [B, C] = their average across exits (:233, what the reference returns for multi-output models).
    """
    from .utils import _MasksemblesBase
    mods = [m for m in model.modules() if isinstance(m, _MasksemblesBase)]
    if not mods:
        raise ValueError("masksembles_batched_predict: the model has no Masksembles layers")
    n = int(mods[0].n)
    if any(int(m.n) != n for m in mods):
        raise ValueError("Masksembles layers with different n in one network")
    saved = [int(m.cnt) for m in mods]
    for m in mods:
        m.cnt = 0
    try:
        r = mc_predict(model, x, n, dtype=dtype)
        per_exit = r.mean_probs.clone()
    finally:
        for m, c in zip(mods, saved):
            m.cnt = c
    return {"per_exit": per_exit, "prediction": per_exit.mean(0), "n": n}
