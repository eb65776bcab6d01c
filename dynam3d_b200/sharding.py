"""Multi-GPU plumbing: episodes are independent, so the hot path shards by episode with NO data-path collective
(the reference splits evaluation trajectories the same way: `traj[rank::world]`, base_il_trainer.py:770).

The only collective is an optional all-gather of per-episode results (next-action logits, or the padded 3D-token memory) for
a consumer that wants the whole batch on every rank; over NVLink 5 / NVSwitch it is a single ncclAllGather.
"""
import torch
import torch.distributed as dist


def shard_episodes(n_episodes, rank, world):
    """Episode ids owned by `rank` (round-robin like the reference's `traj[rank::world]`)."""
    return list(range(rank, n_episodes, world))


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allgather_last_logits(logits):
    """[E, vocab] per rank -> [world*E, vocab] on every rank (rank-major)."""
    w = _world()
    if w == 1:
        return logits
    out = torch.empty((w * logits.shape[0],) + tuple(logits.shape[1:]), device=logits.device, dtype=logits.dtype)
    if logits.is_cuda:
        dist.all_gather_into_tensor(out, logits.contiguous())
    else:
        parts = [torch.empty_like(logits) for _ in range(w)]
        dist.all_gather(parts, logits.contiguous())
        out = torch.cat(parts, 0)
    return out


class LogitsGather:
    """The per-step all-gather of the batch's next-action logits, OFF the critical path: episodes are independent (the reference splits
    trajectories by rank, base_il_trainer.py:770), so the gather of step i only has to be complete when somebody reads it.  `post()` issues
    the collective asynchronously (NCCL: on its own stream behind the producing kernels; gloo: a background thread) and returns at once;
    at most `depth` gathers are in flight (older ones are waited first, bounding memory); `drain()` waits for all and returns the results in
    order.  Without a process group every call degenerates to the identity."""

    def __init__(self, depth=2):
        self.depth, self.pending, self.done = depth, [], []

    def post(self, logits):
        w = _world()
        if w == 1:
            self.done.append(logits)
            return
        out = torch.empty((w * logits.shape[0],) + tuple(logits.shape[1:]), device=logits.device, dtype=logits.dtype)
        src = logits.contiguous()
        if logits.is_cuda:
            work = dist.all_gather_into_tensor(out, src, async_op=True)
        else:
            parts = list(out.chunk(w, 0))
            work = dist.all_gather(parts, src, async_op=True)
        self.pending.append((work, out, src))
        while len(self.pending) > self.depth:
            self._wait_one()

    def _wait_one(self):
        work, out, _ = self.pending.pop(0)
        work.wait()
        self.done.append(out)

    def drain(self):
        while self.pending:
            self._wait_one()
        res, self.done = self.done, []
        return res


def allgather_token_memory(tokens, max_tokens):
    """tokens: list (one per local episode) of [n_i, width] tensors.  Returns (padded [world*E, max_tokens, width], counts [world*E])
    on every rank -- the 'single NCCL all-gather for batched 3D-token memory' of the north star."""
    E = len(tokens)
    width = tokens[0].shape[1]
    dev, dt = tokens[0].device, tokens[0].dtype
    pad = torch.zeros((E, max_tokens, width), device=dev, dtype=dt)
    cnt = torch.zeros((E,), device=dev, dtype=torch.int32)
    for i, t in enumerate(tokens):
        n = min(t.shape[0], max_tokens)
        pad[i, :n].copy_(t[:n])
        cnt[i] = n
    w = _world()
    if w == 1:
        return pad, cnt
    if pad.is_cuda:
        out = torch.empty((w * E, max_tokens, width), device=dev, dtype=dt)
        oc = torch.empty((w * E,), device=dev, dtype=torch.int32)
        dist.all_gather_into_tensor(out, pad)
        dist.all_gather_into_tensor(oc, cnt)
        return out, oc
    parts = [torch.empty_like(pad) for _ in range(w)]
    pc = [torch.empty_like(cnt) for _ in range(w)]
    dist.all_gather(parts, pad)
    dist.all_gather(pc, cnt)
    return torch.cat(parts, 0), torch.cat(pc, 0)
