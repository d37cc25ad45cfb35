import torch
import torch.nn as nn
import torch.nn.functional as F


class GELU(nn.Module):
    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, hidden_states):
        return F.gelu(self.proj(hidden_states), approximate=self.approximate)


class LinearActivation(nn.Module):
    def __init__(self, dim_in, dim_out, bias=True, activation="silu"):
        super().__init__()
        assert activation == "silu"
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.activation = nn.SiLU()

    def forward(self, hidden_states):
        return self.activation(self.proj(hidden_states))


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False,
                 inner_dim=None, bias=True):
        super().__init__()
        inner_dim = inner_dim or int(dim * mult)
        dim_out = dim_out or dim
        if activation_fn == "gelu-approximate":
            act = GELU(dim, inner_dim, approximate="tanh", bias=bias)
        elif activation_fn == "gelu":
            act = GELU(dim, inner_dim, bias=bias)
        elif activation_fn == "linear-silu":
            act = LinearActivation(dim, inner_dim, bias=bias, activation="silu")
        else:
            raise NotImplementedError(activation_fn)
        self.net = nn.ModuleList([act, nn.Dropout(dropout), nn.Linear(inner_dim, dim_out, bias=bias)])

    def forward(self, hidden_states, *args, **kwargs):
        for m in self.net:
            hidden_states = m(hidden_states)
        return hidden_states


class AttentionModuleMixin:
    fused_projections = False  # upstream class attribute (fuse_projections() flips it; never called on this path)

    def set_processor(self, processor):
        self.processor = processor if processor is not None else self._default_processor_cls()

    def get_processor(self):
        return self.processor


class Attention(nn.Module):
    """Only what tests/components/test_attention_list_clear.py of the reference touches."""

    def __init__(self, query_dim, heads=8, dim_head=64, bias=False, **kw):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.norm_q = None
        self.norm_k = None
        self.add_k_proj = None
        self.add_v_proj = None


class AttentionMixin:
    """Processor get/set helpers of upstream; the hot path never calls them."""
    pass
