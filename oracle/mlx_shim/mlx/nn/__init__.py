"""NumPy restatement of the `mlx.nn` layers the reference instantiates (TEST INFRASTRUCTURE ONLY).

Published MLX semantics restated here: `Linear` computes x @ W.T + b with W [out,in];
`LayerNorm` uses the biased variance with eps inside the square root; `Embedding` is a row
gather; `Conv1d` takes channels-last input [N,L,C] and weight [out,K,in/groups]; `gelu` is the
exact erf form; `ELU(alpha)` is x if x>0 else alpha*(exp(x)-1); `SiLU` is x*sigmoid(x).
Initial parameter values are irrelevant: gen_golden.py always loads a checkpoint.
"""

from __future__ import annotations

import numpy as np

from .. import core as mx


class Module:
    """Attribute-tree module (MLX's is a dict subclass; only the calls the reference makes)."""

    def named_modules(self):
        out = []

        def visit(prefix, obj):
            if isinstance(obj, Module):
                out.append((prefix, obj))
                for k, v in vars(obj).items():
                    if k.startswith("_"):
                        continue
                    visit(f"{prefix}.{k}" if prefix else k, v)
            elif isinstance(obj, (list, tuple)):
                for i, v in enumerate(obj):
                    visit(f"{prefix}.{i}" if prefix else str(i), v)
            elif isinstance(obj, dict):
                for k, v in obj.items():
                    visit(f"{prefix}.{k}" if prefix else str(k), v)

        visit("", self)
        return out

    def parameters(self):
        out = {}
        for name, mod in self.named_modules():
            for k, v in vars(mod).items():
                if isinstance(v, mx.array):
                    out[f"{name}.{k}" if name else k] = v
        return out

    def __call__(self, *a, **k):
        raise NotImplementedError


class Identity(Module):
    def __call__(self, x):
        return x


class Linear(Module):
    def __init__(self, input_dims, output_dims, bias=True):
        self.weight = mx.zeros((output_dims, input_dims))
        if bias:
            self.bias = mx.zeros((output_dims,))

    def __call__(self, x):
        y = mx.matmul(x, self.weight.T)
        if hasattr(self, "bias"):
            y = y + self.bias
        return y


class Embedding(Module):
    def __init__(self, num_embeddings, dims):
        self.weight = mx.zeros((num_embeddings, dims))

    def __call__(self, x):
        return self.weight[x]


class LayerNorm(Module):
    def __init__(self, dims, eps=1e-5, affine=True, bias=True):
        self.eps = eps
        self.dims = dims
        if affine:
            self.weight = mx.ones((dims,))
            if bias:
                self.bias = mx.zeros((dims,))

    def __call__(self, x):
        mean = mx.mean(x, axis=-1, keepdims=True)
        var = mx.var(x, axis=-1, keepdims=True)
        y = (x - mean) * mx.rsqrt(var + self.eps)
        if hasattr(self, "weight"):
            y = y * self.weight
        if hasattr(self, "bias"):
            y = y + self.bias
        return y


class Conv1d(Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True):
        self.weight = mx.zeros((out_channels, kernel_size, in_channels // groups))
        if bias:
            self.bias = mx.zeros((out_channels,))
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.groups = groups

    def __call__(self, x):
        y = mx.conv1d(x, self.weight, self.stride, self.padding, self.dilation, self.groups)
        if hasattr(self, "bias"):
            y = y + self.bias
        return y


class ConvTranspose1d(Module):  # only used by the reference in an isinstance() test
    pass


class ELU(Module):
    def __init__(self, alpha=1.0):
        self._alpha = alpha

    def __call__(self, x):
        r = np.asarray(x)
        return mx.array(np.where(r > 0, r, self._alpha * (np.exp(np.minimum(r, 0)) - 1)).astype(r.dtype))


class SiLU(Module):
    def __call__(self, x):
        return x * mx.sigmoid(x)


def gelu(x):
    return x * (1 + mx.erf(x / np.float32(np.sqrt(2.0)))) / 2
