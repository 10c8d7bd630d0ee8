"""TEST INFRASTRUCTURE (never imported by the product path): row-subsample parity of the CUDA `log_prob` against the CPU oracle at
batch sizes the oracle cannot evaluate whole (the B = 8192 / 131072 batches bench.py measures).

The CUDA forward of the WHOLE batch is run eagerly under `rng.record_draws()`, which keeps every random draw (dequantisation,
Augment, encoder noise) exactly as torch's generator produced it, in the reference's draw order.  The oracle
(`flow_oracle.log_prob`, the restatement of reference layers/flowsequential.py:18-30 pinned by tests/golden) then evaluates a row
subsample with the matching rows of those draws.  Every sample of the path is independent of the others, so rows must agree
within the parity gates of SURVEY §8(d)."""
import torch

from contextflow_b200 import rng
from oracle import flow_oracle as O

L_RTOL, L_ATOL = 1e-4, 1e-3          # log-prob gate


def spread_rows(B, n):
    """n row indices spread over [0, B): both ends, CTA / tile boundaries (multiples of 64 and 128 +- 1) and a stride in between."""
    pick = {0, B - 1}
    for edge in (63, 64, 65, 127, 128, 129, 255, 256, 2399, 2400, 4095, 4096):
        if edge < B:
            pick.add(edge); pick.add(B - 1 - edge)
    step = max(1, B // max(1, n - len(pick)))
    pick.update(range(step // 2, B, step))
    return torch.tensor(sorted(pick)[: max(n, len(pick))], dtype=torch.int64)


def recorded_log_prob(model, x, ctx, seed):
    """(log-probs of the eager CUDA forward under torch.manual_seed(seed), DrawRecorder with its draws)."""
    with torch.no_grad(), rng.record_draws() as rec:
        torch.manual_seed(seed)
        logp = model.log_prob_eager(x, ctx)
    return logp, rec


def rows_parity(model, conf, x, ctx, seed, n_rows=256, logp=None, rec=None):
    """dict(rows, max_tol_ratio, max_abs_err, ok): CUDA log_prob rows vs oracle rows; tol-ratio = |err| / (atol + rtol |ref|) <= 1."""
    if logp is None:
        logp, rec = recorded_log_prob(model, x, ctx, seed)
    B = x.shape[0]
    idx = spread_rows(B, n_rows)
    stack = O.build_stack(conf['cfg'], conf['data_size'], conf['mixtures'], conf['contexts'])
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    O.DEVICE = 'cpu'
    with torch.no_grad():
        ref = O.log_prob(stack, state, x.cpu()[idx], None if ctx is None else ctx.cpu()[idx], rec.rows(idx))
    got = logp.detach().cpu()[idx]
    err = (got - ref).abs()
    ratio = (err / (L_ATOL + L_RTOL * ref.abs())).max().item()
    bpd = (O.bits_per_dim(got, conf['data_size']) - O.bits_per_dim(ref, conf['data_size'])).abs().max().item()
    return dict(rows=int(idx.numel()), batch=int(B), max_tol_ratio=ratio, max_abs_err=err.max().item(), max_bpd_err=bpd,
                gate=f'rtol {L_RTOL} + atol {L_ATOL} on log-prob, bpd within 1e-3', ok=bool(ratio <= 1.0 and bpd <= 1e-3))
