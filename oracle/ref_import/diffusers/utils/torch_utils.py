import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """upstream semantics: draw on the generator's device (CPU generator -> CPU draw), then move."""
    device = device or torch.device("cpu")
    gen_device = generator.device if generator is not None else device
    return torch.randn(shape, generator=generator, device=gen_device, dtype=dtype).to(device)


def maybe_allow_in_graph(cls):
    return cls
