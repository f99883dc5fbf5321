"""One process per GPU over torch.distributed (NCCL on the GPU box, gloo in CPU tests).

The reference has no distributed code at all: it shards by hand with `--batch_nums/--batch_index`
(contiguous slices of the 400 loader steps, reference image_main.py:18-19, 61-63, 83) and one python
process per GPU.  Here the same partition is a function of (rank, world), and the two places where
the attack has a real exchange step get a collective:

  * clips / frames are independent units (per-frame loss, BN in eval, element-wise Adam), so I2V,
    ENS-I2V, BIM, FGSM, MI shard with NO data-path collective — `clip_shard`.
  * AENS data-parallel over the frames of one call: prev[l] = sum_n cos[l,n] runs over all frames
    (TPAMI_attack.py:296) -> all-reduce of the [L, N_local] row sums, i.e. L floats per step.
  * ensemble with one backbone per GPU: dcost/dtrue_image = sum over models -> all-reduce(SUM) of the
    [N,3,H,W] gradient once per step, and an all-gather of the per-layer cosines for K2.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for WORLD_SIZE=1)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout for results (NCCL prints its version banner there)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def contiguous_shard(n_items, rank, world):
    """[lo, hi) of `n_items` for `rank`: the reference's --batch_nums/--batch_index split
    (image_main.py:61-63: `range(index*per, (index+1)*per)`), generalised to a remainder."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def clip_shard(n_clips, rank, world, mode="round_robin"):
    """Clip indices handled by `rank`.  round_robin balances a sweep whose clips finish at different
    times; 'contiguous' reproduces the reference's slices."""
    if mode == "round_robin":
        return list(range(rank, n_clips, world))
    lo, hi = contiguous_shard(n_clips, rank, world)
    return list(range(lo, hi))


def ensemble_placement(model_names, rank, world):
    """Which ensemble members `rank` computes and which replica group it belongs to.
    world >= M: one backbone per GPU, world // M data-parallel replicas (ranks beyond M*replicas idle).
    world <  M: members are dealt round-robin."""
    M = len(model_names)
    if world >= M:
        replicas = world // M
        if rank >= replicas * M:
            return [], None
        return [rank % M], rank // M
    return list(range(rank, M, world)), 0


class EnsemblePlan:
    """One backbone per GPU (BASELINE.json configs[2]): which ensemble members this rank computes, where their
    hooked layers sit in the ensemble-wide [L, N] cosine table, and the process group the per-step exchange
    (sum of dcost/dtrue_image, fill-in of the peers' cosine rows) runs over.

    world >= M: rank r computes member r mod M in replica r // M (M ranks per replica, each replica a separate
    group that attacks its own clips); world < M: members are dealt round-robin and the group is the world.
    Every rank must construct the plan (dist.new_group is collective)."""

    def __init__(self, model_names, layers_per_model, rank=None, world=None):
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank, self.world = rank, world
        M = len(model_names)
        if world > M and world % M != 0:
            # a rank outside every replica group would run the attack with no backbone (zero gradient, no collective)
            # and hand back a clean clip as "adversarial": refuse the layout instead
            raise ValueError("one-backbone-per-GPU placement needs a world size that is a multiple of the %d ensemble "
                             "members (or smaller than it), got %d" % (M, world))
        self.members, self.replica = ensemble_placement(model_names, rank, world)
        offsets, off = [], 0
        for n in layers_per_model:
            offsets.append(off)
            off += int(n)
        self.n_layers_total = off
        self.layer_offsets = [offsets[m] for m in self.members]
        self.replicas = max(1, world // M) if world >= M else 1
        self.group = None
        if dist.is_initialized() and world > 1:
            if world >= M:
                for rep in range(self.replicas):                  # collective: same order on every rank
                    g = dist.new_group(list(range(rep * M, (rep + 1) * M)))
                    if rep == self.replica:
                        self.group = g
            else:
                self.group = dist.group.WORLD
        self.active = self.replica is not None

    def hook(self):
        on = self.group is not None
        return ReduceHook(self.group, sum_grad=on, sum_cos=on)


class ReduceHook:
    """Collectives of the frame-sharded AENS loop and of the one-backbone-per-GPU ensemble.

    grad(g)      : all-reduce(SUM) of dcost/dtrue_image across the ensemble group (no-op otherwise)
    cos_rows(cos): for an ensemble group, all-reduce(SUM) fills in the rows computed by peers (each
                   rank writes only its own layers' rows, the others stay 0)
    """

    def __init__(self, group=None, sum_grad=False, sum_cos=False):
        self.group = group
        self.sum_grad = sum_grad
        self.sum_cos = sum_cos

    def grad(self, g):
        if self.sum_grad:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)

    def cos_rows(self, cos):
        if self.sum_cos:
            dist.all_reduce(cos, op=dist.ReduceOp.SUM, group=self.group)


def max_over_ranks(value, device=None):
    """MAX of a python float across ranks (timing rule: a multi-GPU step takes as long as its slowest rank)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            dist.barrier(device_ids=[torch.cuda.current_device()])
        else:
            dist.barrier()
