"""FSDP weight sharding (reference: wan/distributed/fsdp.py) is OUT OF SCOPE: the 14B bf16 weights (28.6 GB) are
replicated on every 180 GB B200.  The symbol exists because text2video imports it."""


def shard_model(model, device_id, **kwargs):
    raise NotImplementedError("FSDP sharding is not part of the B200 hot path: weights are replicated per GPU "
                              "(run without --dit_fsdp / --t5_fsdp)")
