import torch.distributed as dist

_STATE = {"sp": None}


class _SPGroup:
    """What `get_sp_group()` returns: world/rank, all_gather(x, dim) (xdit_context_parallel.py:148), and the
    UlyssesGroup used by the native path."""

    def __init__(self, ulysses_degree, ring_degree):
        from wan.distributed.ulysses import UlyssesGroup
        if ring_degree != 1:
            raise NotImplementedError("ring attention is out of scope (NVSwitch is uniform all-pairs; use "
                                      "--ulysses_size N --ring_size 1)")
        self.ulysses = UlyssesGroup(None)
        self.world_size = self.ulysses.world
        self.rank_in_group = self.ulysses.rank
        if ulysses_degree != self.world_size:
            raise ValueError("ulysses_degree must equal the world size")

    def all_gather(self, x, dim=0):
        import torch
        parts = [torch.empty_like(x) for _ in range(self.world_size)]
        dist.all_gather(parts, x.contiguous())
        return torch.cat(parts, dim=dim)


def init_distributed_environment(rank=None, world_size=None, **kwargs):
    if not dist.is_initialized():
        raise RuntimeError("call torch.distributed.init_process_group first (generate.py:203-207)")


def initialize_model_parallel(sequence_parallel_degree=1, ring_degree=1, ulysses_degree=1, **kwargs):
    _STATE["sp"] = _SPGroup(ulysses_degree, ring_degree)


def get_sp_group():
    if _STATE["sp"] is None:
        raise RuntimeError("initialize_model_parallel has not been called")
    return _STATE["sp"]


def get_sequence_parallel_world_size():
    return get_sp_group().world_size


def get_sequence_parallel_rank():
    return get_sp_group().rank_in_group


def get_world_group():
    return get_sp_group()
