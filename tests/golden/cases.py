"""Golden-fixture case table shared by make_golden.py (reference side) and the tests (oracle / CUDA side)."""
from contextflow_b200.synth import CONFIGS, variant

CASES = {
    # the four BASELINE.json configurations, full architecture, small batch
    'cfg1': dict(conf=CONFIGS['cfg1'], B=4),
    'cfg2': dict(conf=CONFIGS['cfg2'], B=4),
    'cfg3': dict(conf=CONFIGS['cfg3'], B=3),
    'cfg4': dict(conf=CONFIGS['cfg4'], B=5),
    # ActNorm data-dependent initialisation from the first batch (actnorm.py:28-35,46,53)
    'cfg1_init': dict(conf=CONFIGS['cfg1'], B=6, fresh_actnorm=True),
    'cfg4_init': dict(conf=CONFIGS['cfg4'], B=7, fresh_actnorm=True),
    'cfg2_init': dict(conf=variant('cfg2', num_blocks=2, block_size=1), B=5, fresh_actnorm=True),
    # encoder / embedding variants, reduced depth
    'mnist_onehot_uniform': dict(conf=variant('cfg1', generalist=False, contextflow=True, enc_emb='onehot', enc_type='uniform',
                                              num_blocks=2, block_size=1, contexts=[8, 3]), B=3),
    'mnist_eye_uniform': dict(conf=variant('cfg1', generalist=False, contextflow=True, enc_emb='eye', enc_type='uniform',
                                           num_blocks=1, block_size=2, contexts=[64]), B=3),
    'mnist_eye_vardeq2': dict(conf=variant('cfg1', generalist=False, contextflow=True, enc_emb='eye', enc_type='vardeq',
                                           num_blocks=1, block_size=1, contexts=[7, 5]), B=3),
    'mnist_embed_probsample': dict(conf=variant('cfg1', generalist=False, contextflow=True, enc_emb='embed', enc_type='probsample',
                                                num_blocks=1, block_size=1, contexts=[6, 4]), B=3),
    'mnist_embed_eyesample': dict(conf=variant('cfg1', generalist=False, contextflow=True, enc_emb='embed', enc_type='eyesample',
                                               num_blocks=1, block_size=1, contexts=[6]), B=3),
    # (onehot|eye) x eyesample raise inside the reference itself (int64 context into nn.Linear), so they are not cases
    'atm_argmax2': dict(conf=variant('cfg3', num_blocks=1, block_size=2, contexts=[9, 5], data_size=(6, 16, 1)), B=4),
    # conventional (concatenated) context conditioning: contextflow=False specialists (coupling.py:47, conv1x1.py:46-49)
    'cifar_conventional': dict(conf=variant('cfg2', contextflow=False, num_blocks=2, block_size=1), B=3),
    'smap_conventional': dict(conf=variant('cfg4', generalist=False, contextflow=False, enc_emb='onehot', enc_type='vardeq',
                                           num_blocks=1, block_size=2, contexts=[4, 2]), B=3),
    # time-series conv coupling (3,1) reflect kernels; MSL-shaped windows
    'msl_conv': dict(conf=variant('cfg4', dataset='msl', coupling='conv', data_size=(55, 8, 1), contexts=[27],
                                  num_blocks=1, block_size=2), B=3),
    # SMD-shaped windows (38 x 8 x 1, model.py:216-218), ViT generalist
    'smd_trans': dict(conf=variant('cfg4', dataset='smd', data_size=(38, 8, 1), contexts=[28]), B=4),
    # MNIST at the literal 28x28 shape (BASELINE configs[0])
    'mnist28': dict(conf=variant('cfg1', data_size=(1, 28, 28)), B=2),
    # generalist conv coupling on MSL-shaped windows ((3,1) kernels, 56 channels after Augment, M = 2): training-direction case
    # --coupling maf (MaskedCoupling over masked residual conv blocks), generalists: images (3x3) and MSL-shaped windows (3x1)
    'mnist_maf': dict(conf=variant('cfg1', coupling='maf', num_blocks=2, block_size=1), B=3),
    'msl_maf': dict(conf=variant('cfg4', dataset='msl', coupling='maf', data_size=(55, 8, 1), contexts=[27], mixtures=2,
                                 num_blocks=1, block_size=2), B=4),
    # --coupling maf --contextflow specialists: CN = MaskedResidualBlockLinear (ar.py:28, masked_linear.py:104-128), executable in the
    # reference only when the encoder width is 1 (eye over two classes) or 2 x channels (onehot over 16 classes at 8 channels)
    'mnist_maf_eye2': dict(conf=variant('cfg1', coupling='maf', generalist=False, contextflow=True, enc_emb='eye', enc_type='uniform',
                                        num_blocks=1, block_size=2, contexts=[2]), B=4),
    'mnist_maf_onehot16': dict(conf=variant('cfg1', coupling='maf', generalist=False, contextflow=True, enc_emb='onehot', enc_type='uniform',
                                            num_blocks=1, block_size=1, contexts=[16]), B=5),
    # ATM-shaped generalist with the ViT conditioner and PermuteAxes (training-direction case for TransCoupling)
    'atm_gen': dict(conf=variant('cfg3', generalist=True, contextflow=False, num_blocks=1, block_size=2, contexts=[9], data_size=(6, 16, 1)), B=4),
    # --contextflow specialists with the reference's DEFAULT encoder (--enc-emb onehot --enc-type uniform, config.py:18-19): the
    # training-direction cases of the context-conditioned layers (conv with split priors and two context features; ViT)
    'cifar_onehot_uniform': dict(conf=variant('cfg2', enc_type='uniform', num_blocks=2, block_size=1), B=3),
    # the BASELINE cfg2 encoder (onehot + vardeq: trainable inner flows), reduced depth, ActNorms initialised
    'cifar_vardeq': dict(conf=variant('cfg2', num_blocks=2, block_size=1), B=3),
    'atm_onehot_uniform': dict(conf=variant('cfg3', enc_emb='onehot', enc_type='uniform', num_blocks=2, block_size=1, contexts=[9, 5],
                                            data_size=(6, 16, 1)), B=4),
    # the CIFAR generalist (stage one of the paper's workflow): conv stack WITH split priors, no context
    'cifar_gen': dict(conf=variant('cfg2', generalist=True, contextflow=False, num_blocks=2, block_size=1), B=3),
    'msl_conv_gen': dict(conf=variant('cfg4', dataset='msl', coupling='conv', data_size=(55, 8, 1), contexts=[27], mixtures=2,
                                      num_blocks=1, block_size=2), B=6),
}

# ---- inverse direction (make_golden_inverse.py) ----
# models whose every layer has an executable .reverse in the reference: generalists without split priors
INVERSE_CHAIN = ['cfg1', 'cfg4', 'mnist28']
# specialists: per-layer Coupling / TransCoupling .reverse with a context encoder (the other specialist reverses raise in the reference)
INVERSE_COUPLING = ['cfg2', 'cifar_conventional', 'smap_conventional', 'atm_argmax2', 'msl_conv']

# ---- loss / score epilogue ----
SCORE_CASES = {
    'ad_m2': dict(B=37, M=2, size=(25, 8, 1), spread=40.0, shift=-300.0, weight=[0.3, 1.7]),        # anomaly detection, criterion
    'ad_m1': dict(B=9, M=1, size=(25, 8, 1), spread=40.0, shift=-300.0),                            # unsupervised (M = 1)
    'cl_m10': dict(B=64, M=10, size=(3, 32, 32), spread=300.0, shift=-9000.0, weight=[1.0] * 10),   # classification
    'cl_nan': dict(B=12, M=10, size=(1, 28, 28), spread=100.0, shift=-2000.0, nan=True),            # NaN -> 0 replacement
}

# ---- training direction (make_golden_training.py): context-free conv stacks; loss of experiment_ad.py:204-208 ----
TRAINING_CASES = {
    'cfg1': dict(alpha=1e-2, criterion=True, weight=[1.0, 0.5, 2.0, 1.0, 1.0, 1.5, 1.0, 0.7, 1.0, 1.2]),
    'mnist28': dict(alpha=1e-2, criterion=True, weight=None),
    'msl_conv_gen': dict(alpha=1e2, criterion=False, weight=None),
    'cifar_gen': dict(alpha=1e-3, criterion=True, weight=None),           # experiment_cl.py:56 alpha with a criterion
    'cfg4': dict(alpha=1e2, criterion=False, weight=None),                # SMAP generalist, ViT conditioner, M = 1 (experiment_ad.py:61)
    'atm_gen': dict(alpha=1e-2, criterion=True, weight=[0.4, 1.6]),       # ViT conditioner with 8 / 3 tokens, PermuteAxes, M = 2
    # --contextflow specialists, default (parameter-free) encoders: CN networks and the priors' embedding tables train
    'mnist_onehot_uniform': dict(alpha=1e-2, criterion=True, weight=None),
    'cifar_onehot_uniform': dict(alpha=1e-3, criterion=True, weight=None),
    'atm_onehot_uniform': dict(alpha=1e-2, criterion=True, weight=[0.4, 1.6]),
    # --contextflow specialist with the BASELINE cfg2 encoder (variational dequantisation): the encoder flows train too
    'cifar_vardeq': dict(alpha=1e-3, criterion=True, weight=None),
    # the other valid (embedding x surjection) pairs: eye + argmax over the ViT stack (BASELINE cfg3's encoder), eye + vardeq,
    # embed + probsample, embed + eyesample
    'atm_argmax2': dict(alpha=1e-2, criterion=True, weight=None),
    'mnist_eye_vardeq2': dict(alpha=1e-2, criterion=True, weight=None),
    'mnist_embed_probsample': dict(alpha=1e-2, criterion=True, weight=None),
    'mnist_embed_eyesample': dict(alpha=1e-2, criterion=True, weight=None),
    # conventional (concatenated-context, contextflow=False) specialists: every layer trains, the context enters by concatenation
    'cifar_conventional': dict(alpha=1e-3, criterion=True, weight=None),
    'smap_conventional': dict(alpha=1e2, criterion=False, weight=None),
    # --coupling maf generalists (masked residual conv blocks)
    'mnist_maf': dict(alpha=1e-2, criterion=True, weight=None),
    'msl_maf': dict(alpha=1e-2, criterion=True, weight=[0.4, 1.6]),
}
