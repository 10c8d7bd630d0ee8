"""Gaussian densities of the path (reference layers/distributions/gaussian.py):
StandardNormal (Augment noise, :10-72), GaussianMixtureDistribution (base + split priors, :118-169),
ConditionalGaussianDistribution (encoder base, :234-270)."""
import torch
import torch.nn as nn

from ... import ops, rng, training
from ..context import ContextPlan
from ..flowlayer import PackCache, inference_only


__all__ = ['StandardNormal', 'GaussianMixtureDistribution', 'ConditionalGaussianDistribution']

class StandardNormal(nn.Module):
    """N(0, I) over `size`; only the context-free form is on the path (model.py:122, :62)."""

    def __init__(self, size, mixtures=1, context_net=None, contextflow=False):
        super().__init__()
        assert mixtures == 1, 'mixtures should be 1 in GaussianDistribution'
        if context_net:
            raise NotImplementedError('the class-conditional StandardNormal variant is not instantiated by create_model')
        self.size = torch.Size(size)
        self.M, self.K = mixtures, 8
        self.register_buffer('buffer', torch.zeros(1))
        self.context_net, self.contextflow = context_net, contextflow

    def forward(self, input, context=None):
        return self.log_prob(input, context)

    def log_prob(self, input, context=None):
        flat = input.reshape(input.shape[0], -1, 1, 1)
        return -ops.augment(None, flat)[1].unsqueeze(-1)

    def draw(self, n_samples, device=None):
        return rng.randn((n_samples, *self.size), self.buffer.device if device is None else device, self.buffer.dtype)

    def sample(self, n_samples, context=None):
        x = self.draw(n_samples)
        return x, self.log_prob(x, context)


class GaussianMixtureDistribution(nn.Module):
    """K-component, M-mixture diagonal GMM; context adds per-sample mean / pre-softplus scale offsets."""

    def __init__(self, size, mixtures=2, components=8, context_net=None, contextflow=False):
        super().__init__()
        self.size = size
        D, H, W = size
        self.D, self.M, self.K = D, mixtures, components
        self.mG = nn.Parameter(torch.randn(self.M, self.K, D, H, W))
        self.sG = nn.Parameter(torch.ones(self.M, self.K, D, H, W))
        self.wG = nn.Parameter(torch.randn(self.M, self.K))
        self.context_net, self.contextflow = context_net, contextflow
        if self.context_net:
            self.C = self.context_net.C
            if self.contextflow:
                for p in (self.mG, self.sG, self.wG):
                    p.requires_grad_(False)
        self._plan = ContextPlan()
        self._tables = PackCache()

    def log_prob(self, input, context=None):
        if not self.context_net and training.wants_grad(input, self.mG, self.sG, self.wG):
            return training.GmmFn.apply(input, self.mG, self.sG, self.wG, self)
        if self.context_net and training.wants_grad(input, *self.context_net.parameters()):
            tables = training._lookup_tables(self)                 # specialist prior: embedding tables train
            if tables is None:
                raise NotImplementedError('training a mixture whose context_net is not the embed + eyesample lookup of create_model '
                                          'has no backward kernel yet')
            if isinstance(context, list):
                context = context[0]
            return training.GmmCtxFn.apply(input, context, self, 0, self.mG, self.sG, self.wG, *tables)
        inference_only(self.mG); inference_only(input)
        H, W = input.shape[2], input.shape[3]
        if isinstance(context, list):
            context = context[0]
        B = input.shape[0]
        if self.context_net:
            from .._encoder_desc import EncoderBatch
            fused = self._plan.fused_for(self.context_net) if self._plan.preset is None else None
            if fused is not None and EncoderBatch.is_lookup(fused) and context.dim() == 2:
                # embed + eyesample (model.py:157,162): offsets are a table lookup with logp_c = 0
                cards = [int(e.num_embeddings) for e in fused.emb._embeddings]
                tables = fused.emb.tables()
                MKD = self.M * self.K * self.D
                n = len(cards)
                if n in (1, 2) and tables[0].shape[1] * n == 2 * MKD:
                    sf = n - 1                                              # 'b (p m k d)': the scale half comes from the last feature
                    soff = 0 if n == 2 else MKD
                    nkeys = cards[0] * (cards[1] if n == 2 else 1)
                    if B >= 32 * nkeys and nkeys <= 4096:                   # context buckets fill the 64-sample tiles at least half
                        tab = self._tables.get('ctx', [self.mG, self.sG, self.wG, tables[sf]],
                                               lambda: ops.gmm_tile_table(self.mG, self.sG, self.wG, tables[sf], soff))
                        if tab is not None:
                            return ops.gmm_tile_logprob(input, tab, self.M, self.K, context, cards, tables[0], 0)
                keep = self._tables.get('ctxtab_ws', [self.mG, self.sG, self.wG, *tables], dict)    # workspaces by batch size, per parameter version
                out = ops.gmm_logprob_ctxtab(input, self.mG, self.sG, self.wG, context, cards, tables, keep=keep)
                if out is not None:
                    return out
            c, logp_c = self._plan.run(self.context_net, context)       # c: 'b (p m k d)'
            return ops.gmm_logprob(input, self.mG, self.sG, self.wG, c, logp_c, float(H * W))
        tab = self._tables.get('plain', [self.mG, self.sG, self.wG], lambda: ops.gmm_tile_table(self.mG, self.sG, self.wG))
        if tab is not None and B > 0:
            return ops.gmm_tile_logprob(input, tab, self.M, self.K)
        return ops.gmm_logprob(input, self.mG, self.sG, self.wG)

    def sample(self, n_samples, context=None):
        """gaussian.py:163-169: draw from the context-free MixtureSameFamily(Categorical(softmax wG), Normal(mG, softplus sG)),
        keep mixture m = 1 (`x[:, 1]`, so M >= 2 is required exactly as in the reference), return (x, log_prob(x, context)).
        Same distribution as the reference, different generator stream: the reference draws all M*K component samples and gathers;
        here one component index per sample (rng.multinomial) and one standard normal per element (rng.randn) are drawn."""
        inference_only(self.mG)
        if self.M < 2:
            raise IndexError('index 1 is out of bounds for dimension 1 with size 1')     # gaussian.py:167 `x[:,1,...]`
        dev = self.mG.device
        probs = self._tables.get('w1', [self.wG], lambda: torch.softmax(self.wG.detach()[1].double(), -1).float())
        comp = rng.multinomial(probs, n_samples)
        eps = rng.randn((n_samples, *self.size), dev, self.mG.dtype)
        x = ops.gmm_sample(self.mG.detach(), self.sG.detach(), comp, eps, m=1)
        return x, self.log_prob(x, context)


class ConditionalGaussianDistribution(nn.Module):
    """q(u | context) = N(mean(context), exp(log_scale(context))): base of the encoder flows."""

    def __init__(self, size, mixtures=1, context_net=None, contextflow=False):
        super().__init__()
        self.size = size
        self.D = size[0]
        assert mixtures == 1, 'mixtures should be 1 in GaussianDistribution'
        self.M = mixtures
        self.context_net, self.contextflow = context_net, contextflow

    def forward(self, x, context=None):
        return self.log_prob(x, context)

    def log_prob(self, x, context=None):
        raise NotImplementedError('only ConditionalGaussianDistribution.sample is on the path (flowsequential.py:61)')

    def sample(self, n_samples, context=None):
        from .._encoder_desc import cond_gauss_sample
        return cond_gauss_sample(self, n_samples, context)
