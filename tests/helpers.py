"""Shared test plumbing: golden loading, oracle state construction, tolerances (SURVEY §8d parity gates)."""
import json, os
import numpy as np
import torch

from contextflow_b200 import synth
from oracle import flow_oracle as O
from tests.golden.cases import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# parity gates: z rtol 1e-4/atol 1e-5; ldj & logp rtol 1e-4 (+atol 1e-3 for near-zero slogdet terms)
Z_RTOL, Z_ATOL = 1e-4, 1e-5
L_RTOL, L_ATOL = 1e-4, 1e-3
BPD_ATOL = 1e-3
BPD_RTOL = 2e-6      # only matters where |bpd| >> 1 (the maf stacks reach ~2e3: one float32 ulp of their log-prob is already 2e-4 bpd)


def load_golden(name):
    g = dict(np.load(os.path.join(GOLD, f'{name}.npz'), allow_pickle=False))
    g['draws'] = json.loads(str(g['draws']))
    g['keys'] = json.loads(str(g['keys']))
    g['layer_types'] = json.loads(str(g['layer_types']))
    return g


def golden_state(g, case):
    """Rebuild the reference state_dict of a golden case from its recorded key->shape table + synth fill."""
    state = {}
    for k, shp in g['keys'].items():
        leaf = k.rsplit('.', 1)[-1]
        if leaf in ('initialized', 'cardinalities'):
            state[k] = torch.zeros(shp, dtype=torch.int64)
        else:
            state[k] = torch.zeros(shp, dtype=torch.float32)
    conf = case['conf']
    # buffers that fill_state leaves alone get their reference values
    for k in state:
        leaf = k.rsplit('.', 1)[-1]
        if leaf == 'temperature':
            state[k].fill_(1.0)
        elif leaf == 'cardinalities':
            state[k].copy_(torch.tensor(conf['contexts']))
    stack = O.build_stack(conf['cfg'], conf['data_size'], conf['mixtures'], conf['contexts'])
    for lay in stack['layers'] + [stack['base']]:
        enc = lay.get('enc')
        if enc is None or enc['num_cats'] is None:
            continue
        pre = f"{lay['key']}.context_net.1"
        if f'{pre}.qbins' in state:
            q = torch.tensor(enc['num_cats'], dtype=torch.float32)
            state[f'{pre}.qbins'].copy_(q); state[f'{pre}.ldj_per_dim'].copy_(-torch.log(q))
    for lay in stack['layers']:
        if lay['op'] == 'maf':                                      # buffers of MaskedConv2d (masked_conv_2d.py:77-78)
            for name in ('conv1', 'conv2', 'conv3'):
                mk = f"{lay['key']}.NN.{name}.mask"
                o, i, kh, kw = state[mk].shape
                state[mk].copy_(O.maf_mask(o, i, kh, kw, lay['C']))
    synth.fill_state(state, case.get('wseed', 'w0'))
    if case.get('fresh_actnorm'):
        for k in state:
            if k.endswith('.initialized'):
                state[k].fill_(0)
    return stack, state


def case_inputs(case):
    return synth.make_inputs(case['conf'], case['B'], case.get('iseed', 'in0'))


def assert_close(a, b, rtol, atol, what=''):
    a = torch.as_tensor(np.asarray(a)).double(); b = torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), f'{what}: max abs err {err.max().item():.3e} (|ref| {b.abs().max().item():.3e}); {int(bad.sum())} of {bad.numel()} outside rtol={rtol} atol={atol}'
