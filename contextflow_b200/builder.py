"""Model construction without the reference on the path: ContextEncoder and create_model with the reference's signatures
(model.py:30-90, :95-163), expressed over this package's layers.  bench.py, smoke() and the GPU tests build their models
here; tests/test_dropin_cpu.py checks (in the build container) that the reference's own create_model, run unmodified over
this package's `layers`, yields the identical module tree and state_dict."""
from __future__ import annotations

import torch.nn as nn

from . import layers as L
from .layers.rtdl.nn._embeddings import CatEmbeddings, EyeEncoder, OneHotEncoder

GMM_COMPONENTS = 8
IMAGE_DATASETS = ('mnist', 'cifar10')
TS_DATASETS = ('atm', 'msl', 'smd', 'smap')
NO_SQUEEZE = ('msl', 'smd', 'smap')

# per-dataset constants of model.main (model.py:173-220): data_size, mixtures, (blocks, block_size), split prior, contexts
DATASETS = {
    'mnist': dict(data_size=(1, 32, 32), mixtures=10, num_blocks=2, block_size=2, split_prior=False, contexts=[64]),
    'cifar10': dict(data_size=(3, 32, 32), mixtures=10, num_blocks=3, block_size=4, split_prior=True, contexts=[15, 5]),
    'atm': dict(data_size=(38, 144, 1), mixtures=2, num_blocks=3, block_size=4, split_prior=True, contexts=[68]),
    'smap': dict(data_size=(25, 8, 1), mixtures=1, num_blocks=2, block_size=4, split_prior=False, contexts=[55]),
    'msl': dict(data_size=(55, 8, 1), mixtures=1, num_blocks=2, block_size=4, split_prior=False, contexts=[27]),
    'smd': dict(data_size=(38, 8, 1), mixtures=1, num_blocks=2, block_size=4, split_prior=False, contexts=[28]),
}


class ContextEncoder(nn.Sequential):
    """context (B, n) int64 -> (c (B, C), logp_c (B,)): categorical embedding followed by a surjective dequantiser."""

    def __init__(self, contexts, enc_emb, enc_type, data_size, init='orthogonal'):
        n = len(contexts)
        num_cats = None
        if enc_emb == 'onehot':
            width, emb, num_cats = sum(contexts), OneHotEncoder(contexts), sum(contexts) * [1]
        elif enc_emb == 'eye':
            emb, num_cats = EyeEncoder(), contexts
            if enc_type == 'argmax':
                width = sum(L.ArgmaxCatDequantization.cats2bits(contexts))
                width += width % 2
            else:
                width = n
        elif enc_emb == 'embed':
            emb, width = CatEmbeddings(contexts, data_size[0], stack=False, init=init), data_size[0] * n
        else:
            raise NotImplementedError('{} is not supported enc-emb!'.format(enc_emb))

        def inner_flow():
            if width % 2:
                raise NotImplementedError('odd encoder widths need the Augment step that the reference mis-shapes (model.py:61-63)')
            steps = []
            for _ in range(2):
                steps += [L.FC((width,)), L.ActNormFC((width,)), L.CouplingFC(width)]
            base = L.ConditionalGaussianDistribution(size=(width,), context_net=CatEmbeddings(contexts, 2 * width // n, stack=False, init='zeros'))
            return L.FlowInvSequential(base, *steps)

        if enc_type == 'eyesample':
            surj = L.EyeSampling()
        elif enc_type == 'probsample':
            surj = L.ProbSampling(inner_flow())
        elif enc_type in ('uniform', 'vardeq', 'argmax'):
            if num_cats is None:
                raise NotImplementedError(f'enc_emb=embed has no category counts for enc_type={enc_type} (undefined in the reference too)')
            if enc_type == 'uniform':
                surj = L.UniformCatDequantization(num_cats=num_cats)
            elif enc_type == 'vardeq':
                surj = L.VariationalCatDequantization(inner_flow(), num_cats=num_cats)
            else:
                surj = L.ArgmaxCatDequantization(inner_flow(), num_cats=num_cats)
        else:
            raise NotImplementedError('{} is not supported enc-type!'.format(enc_type))
        self.C = width
        self.contexts = contexts
        super().__init__(emb, surj)


def create_model(config, data_size=(1, 1, 1), mixtures=1, contexts=[-1]):
    """Same arguments and layer stack as the reference's create_model; `config['dataset']` replaces its global `c.dataset`."""
    dataset = config['dataset']
    alpha = 1e-4
    stack = []
    if dataset in IMAGE_DATASETS:
        stack += [L.Dequantization(L.UniformDistribution(size=data_size)), L.Normalization(translation=0.0, scale=256.0),
                  L.Normalization(translation=alpha, scale=1 / (1 - 2 * alpha)), L.LogitTransform()]
    if not (mixtures == 1 or config['dist'] == 'gauss'):
        raise NotImplementedError('{} is not supported base distribution!'.format(config['dist']))
    ts = dataset in TS_DATASETS
    patch, krn, pad = ((2, 1), (3, 1), (1, 0)) if ts else ((2, 2), (3, 3), (1, 1))
    cf, specialist = config['contextflow'], not config['generalist']

    def ctx_net(d0, emb=None, typ=None, init='orthogonal'):
        if not specialist:
            return None
        return ContextEncoder(contexts, emb or config['enc_emb'], typ or config['enc_type'], (d0,), init=init)

    def prior(sz):
        width = 2 * mixtures * GMM_COMPONENTS * sz[0] // len(contexts)
        return L.GaussianMixtureDistribution(size=sz, mixtures=mixtures, components=GMM_COMPONENTS,
                                             context_net=ctx_net(width, 'embed', 'eyesample', init='zeros'), contextflow=cf)

    sz = tuple(data_size)
    for blk in range(config['num_blocks']):
        if sz[0] % 2:
            stack.append(L.Augment(L.StandardNormal((1, sz[1], sz[2])), 1))
            sz = (sz[0] + 1, sz[1], sz[2])
        if dataset not in NO_SQUEEZE:
            stack.append(L.Squeeze(patch_size=patch))
            sz = (sz[0] * patch[0] * patch[1], sz[1] // patch[0], sz[2] // patch[1])
        for _ in range(config['block_size']):
            stack.append(L.Conv1x1(sz, context_net=ctx_net(sz[0], init='zeros'), contextflow=cf))
            if config['actnorm']:
                stack.append(L.ActNorm(sz, context_net=ctx_net(2 * sz[0]), contextflow=cf))
            if config['coupling'] == 'trans' and sz[1] % patch[0] == 0 and sz[2] % patch[1] == 0:
                stack.append(L.TransCoupling(sz, patch, context_net=ctx_net(sz[0]), contextflow=cf))
            elif config['coupling'] == 'conv':
                stack.append(L.Coupling(sz[0], kernel_size=krn, padding=pad, context_net=ctx_net(sz[0]), contextflow=cf))
            elif config['coupling'] == 'maf':
                stack.append(L.MaskedCoupling(sz[0], kernel_size=krn, padding=pad, context_net=ctx_net(sz[0]), contextflow=cf))
            if dataset == 'atm':
                stack.append(L.PermuteAxes((0, 2, 1, 3)))
                sz = (sz[1], sz[0], sz[2])
        if config['split_prior'] and blk < config['num_blocks'] - 1:
            sz = (sz[0] // 2, sz[1], sz[2])
            stack.append(L.SplitPrior(prior(sz)))
    return L.FlowSequential(prior(sz), *stack)


def build_named(conf: dict):
    """Model for one entry of synth.CONFIGS / synth.variant()."""
    return create_model(conf['cfg'], data_size=conf['data_size'], mixtures=conf['mixtures'], contexts=conf['contexts'])
