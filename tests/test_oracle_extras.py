"""Pin the oracle's restatement of the standalone Sigmoid / Softplus flow layers and of StudentMixtureDistribution.log_prob against
outputs of the unmodified reference classes (tests/golden/extras.npz from make_golden_extras.py).  CPU only."""
import os
import numpy as np
import torch

from contextflow_b200 import synth
from oracle import flow_oracle as O
from tests.golden.make_golden_extras import STUDENT
from tests.helpers import GOLD, assert_close


def gold():
    return dict(np.load(os.path.join(GOLD, 'extras.npz'), allow_pickle=False))


def student_params():
    D, H, W = STUDENT['size']
    M, K = STUDENT['mixtures'], 8
    P = {k: torch.zeros((M, K, D, H, W)) for k in ('mG', 'sG', 'mS', 'sS')}
    P['wG'], P['wS'] = torch.zeros(M, K), torch.zeros(M, K)
    P['vS'] = torch.linspace(1, 10, K).view(1, K, 1, 1, 1).repeat(M, 1, D, H, W)           # student.py:56-58
    synth.fill_state(P, 'student')
    return P


def test_oracle_activation_layers_match_reference():
    g = gold()
    x = torch.from_numpy(g['act_x'])
    for T in (1.0, 2.5):
        z, ldj = O.sigmoid_layer(x, T)
        assert_close(z.numpy(), g[f'sig_z_{T}'], 1e-6, 1e-7, f'sigmoid z T={T}')
        assert_close(ldj.numpy(), g[f'sig_ldj_{T}'], 1e-6, 1e-5, f'sigmoid ldj T={T}')
        assert_close(O.sigmoid_layer_reverse(z, T, 1e-6).numpy(), g[f'sig_rev_{T}'], 1e-6, 1e-6, f'sigmoid reverse T={T}')
    z, ldj = O.softplus_layer(x)
    assert_close(z.numpy(), g['sp_z'], 1e-6, 1e-7, 'softplus z')
    assert_close(ldj.numpy(), g['sp_ldj'], 1e-6, 1e-5, 'softplus ldj')
    assert_close(O.softplus_layer_reverse(z).numpy(), g['sp_rev'], 1e-6, 1e-6, 'softplus reverse')


def test_oracle_student_mixture_matches_reference():
    g = gold()
    P = student_params()
    x = torch.from_numpy(g['stu_x'])
    assert_close(O.student_mixture_log_prob(P, x).numpy(), g['stu_logp'], 1e-5, 1e-4, 'student log_prob')
    P64 = {k: v.double() for k, v in P.items()}
    assert_close(O.student_mixture_log_prob(P64, x.double()).numpy(), g['stu_logp64'], 1e-9, 1e-9, 'student log_prob (float64)')
