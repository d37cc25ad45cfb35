"""Stand-in: context-parallel plan records referenced by class attributes of the Flux2 model (never used on this path)."""


class ContextParallelInput:
    def __init__(self, split_dim=None, expected_dims=None, split_output=False):
        self.split_dim, self.expected_dims, self.split_output = split_dim, expected_dims, split_output


class ContextParallelOutput:
    def __init__(self, gather_dim=None, expected_dims=None):
        self.gather_dim, self.expected_dims = gather_dim, expected_dims
