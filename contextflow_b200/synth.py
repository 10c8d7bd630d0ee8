"""Deterministic synthetic weights, inputs and noise for the flow log-density path.

Everything here is derived from a splitmix64 hash of (tag, element index) using integer
arithmetic only, so the very same numbers come out on any machine (the golden fixtures under
tests/golden/ were produced with these generators while running the *reference* in the build
container; the GPU box regenerates the identical weights/inputs without the reference).

Used by: tests/, tests/golden/make_golden.py, bench.py, __graft_entry__.smoke().
"""
from __future__ import annotations

import zlib
import numpy as np
import torch

_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _mix(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _base(tag: str) -> np.uint64:
    b = tag.encode()
    lo = zlib.crc32(b)
    hi = zlib.crc32(b[::-1] + b'#cfpp')
    return np.uint64((hi << 32) | lo)


def u01(tag: str, n: int, stream: int = 0) -> np.ndarray:
    """n float64 values in [0,1), a pure function of (tag, stream, index)."""
    with np.errstate(over='ignore'):
        idx = np.arange(n, dtype=np.uint64)
        h = _mix(_base(tag) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03) + (idx + np.uint64(1)) * _GOLD)
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def uniform(tag: str, shape, lo=-1.0, hi=1.0, dtype=torch.float32) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    v = lo + (hi - lo) * u01(tag, n)
    return torch.from_numpy(v).reshape(tuple(shape)).to(dtype)


def normal(tag: str, shape, dtype=torch.float32) -> torch.Tensor:
    """Irwin-Hall(12)-6: a portable stand-in for N(0,1) draws (adds only, no libm)."""
    n = int(np.prod(shape)) if len(shape) else 1
    acc = np.zeros(n, dtype=np.float64)
    for s in range(12):
        acc += u01(tag, n, stream=s + 1)
    return torch.from_numpy(acc - 6.0).reshape(tuple(shape)).to(dtype)


def randint(tag: str, shape, high: int) -> torch.Tensor:
    n = int(np.prod(shape)) if len(shape) else 1
    v = np.floor(u01(tag, n) * high).astype(np.int64)
    return torch.from_numpy(np.minimum(v, high - 1)).reshape(tuple(shape))


class NoiseTape:
    """The path's random draws, replayed in call order (RNG contract: SURVEY App. C-7).

    The n-th draw of a forward is `normal/uniform('<name>/<n>', shape)`; reference, oracle and CUDA
    path all see identical numbers as long as they draw in the same order with the same shapes.
    """

    def __init__(self, name: str):
        self.name, self.n, self.log = name, 0, []

    def _tag(self, kind, shape):
        t = f'{self.name}/{self.n}'
        self.log.append((kind, tuple(int(s) for s in shape)))
        self.n += 1
        return t

    def rand(self, shape, device=None, dtype=torch.float32):
        t = uniform(self._tag('rand', shape), shape, 0.0, 1.0, torch.float64).to(torch.float32)
        t = torch.clamp(t, max=float(np.nextafter(np.float32(1.0), np.float32(0.0))))
        return t.to(device=device, dtype=dtype)

    def randn(self, shape, device=None, dtype=torch.float32):
        return normal(self._tag('randn', shape), shape, torch.float32).to(device=device, dtype=dtype)


def _fan_in(shape):
    return int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])


def fill_state(state: dict, seed: str = 'w0') -> dict:
    """Overwrite every learnable tensor of a flow state_dict in place, by key pattern.

    Keys follow the reference's state_dict names, which the product's modules share.
    """
    for key in sorted(state.keys()):
        v = state[key]
        leaf = key.rsplit('.', 1)[-1]
        tag = f'{seed}:{key}'
        shape = tuple(v.shape)
        if leaf == 'initialized':
            v.fill_(1)
        elif leaf in ('qbins', 'ldj_per_dim', 'cardinalities', 'temperature', 'translation', 'scale', 'empty', 'buffer', 'mask', 'degrees'):
            continue
        elif leaf == 'NN' and v.dim() == 2:                      # invertible 1x1 / FC matrix
            d = shape[0]
            v.copy_(torch.eye(d) + uniform(tag, shape) * (0.25 / np.sqrt(d)))
        elif leaf in ('NN_t', 'NN_logs'):
            v.copy_(uniform(tag, shape) * 0.2)
        elif leaf in ('mG', 'mS'):
            v.copy_(uniform(tag, shape))
        elif leaf in ('sG', 'sS'):
            v.copy_(1.0 + 0.3 * uniform(tag, shape))
        elif leaf in ('wG', 'wS'):
            v.copy_(uniform(tag, shape))
        elif leaf == 'vS':                                       # Student-t degrees of freedom (pre-softplus): keeps the linspace(1, 10) initialisation
            continue
        elif '_embeddings.' in key:
            v.copy_(uniform(tag, shape) * 0.1)
        elif leaf == 'weight' and v.dim() == 1:                  # LayerNorm gain (the ViT's output norm is kept small)
            g = 0.25 if key.endswith('transformer.norm.weight') else 1.0
            v.copy_(g * (1.0 + 0.1 * uniform(tag, shape)))
        elif leaf == 'weight':
            scale = 1.0 / np.sqrt(_fan_in(shape))
            if '.CN.' in key or key.startswith('CN.'):
                scale *= 0.5
            if key.endswith('CN.weight') and (key[:-len('CN.weight')] + 'NN') in state:  # Conv1x1's D*D context matrix
                scale *= 0.5 / float(shape[0]) ** 0.25
            if key.endswith('NN.4.weight') or key.endswith('CN.4.weight'):   # conditioner output layer: keep |h| ~ 0.3
                scale *= 0.4
            if '.NN.conv' in key:                                            # masked residual block (--coupling maf): h rides on the identity x
                scale *= 0.15 if key.endswith('conv3.weight') else 0.5
            v.copy_(uniform(tag, shape) * scale)
        elif leaf == 'bias':
            v.copy_(uniform(tag, shape) * (0.03 if key.endswith('transformer.norm.bias') else 0.1))
        else:
            raise KeyError(f'fill_state: no rule for {key} {shape}')
    return state


# ----- the BASELINE.json configurations -------------------------------------
CONFIGS = {
    'cfg1': dict(cfg=dict(dataset='mnist', contextflow=False, generalist=True, enc_emb='eye', enc_type='uniform',
                          num_blocks=2, block_size=2, actnorm=True, split_prior=False, coupling='conv', dist='gauss'),
                 data_size=(1, 32, 32), mixtures=10, contexts=[64], image=True),
    'cfg2': dict(cfg=dict(dataset='cifar10', contextflow=True, generalist=False, enc_emb='onehot', enc_type='vardeq',
                          num_blocks=3, block_size=4, actnorm=True, split_prior=True, coupling='conv', dist='gauss'),
                 data_size=(3, 32, 32), mixtures=10, contexts=[15, 5], image=True),
    'cfg3': dict(cfg=dict(dataset='atm', contextflow=True, generalist=False, enc_emb='eye', enc_type='argmax',
                          num_blocks=3, block_size=4, actnorm=True, split_prior=True, coupling='trans', dist='gauss'),
                 data_size=(38, 144, 1), mixtures=2, contexts=[68], image=False),
    'cfg4': dict(cfg=dict(dataset='smap', contextflow=False, generalist=True, enc_emb='eye', enc_type='uniform',
                          num_blocks=2, block_size=4, actnorm=True, split_prior=False, coupling='trans', dist='gauss'),
                 data_size=(25, 8, 1), mixtures=1, contexts=[55], image=False),
}


def variant(base: str, **over) -> dict:
    import copy
    c = copy.deepcopy(CONFIGS[base])
    for k, v in over.items():
        if k in c['cfg']:
            c['cfg'][k] = v
        else:
            c[k] = v
    return c


def make_inputs(conf: dict, B: int, tag: str = 'in0'):
    """x (B,C,H,W) float32 and context (B,n_ctx) int64 of the dataset's shape."""
    C, H, W = conf['data_size']
    if conf['image']:
        x = randint(f'{tag}:x', (B, C, H, W), 256).to(torch.float32)
    else:
        x = uniform(f'{tag}:x', (B, C, H, W), 0.0, 1.0)
    ctx = torch.stack([randint(f'{tag}:c{i}', (B,), card) for i, card in enumerate(conf['contexts'])], 1)
    return x, ctx
