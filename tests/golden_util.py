"""Helpers shared by the oracle tests (CPU) and the parity tests (GPU)."""
import os

import numpy as np
import torch

from oracle import nerf_oracle as orc
from simplenerf_b200 import synthetic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RENDER_CASES = {
    'render_llff_simplenerf.npz': ('simplenerf', True, 'llff'),
    'render_llff_simplenerf_dense.npz': ('simplenerf', True, 'llff'),
    'render_re10k_vanilla_dense.npz': ('vanilla', True, 're10k'),
    'render_nondc_vanilla_dense.npz': ('vanilla', False, 'llff'),
    'render_llff_fineaug_dense.npz': ('simplenerf_fineaug', True, 'llff'),       # fine-level augmentation MLPs (row N4)
}


def load(name):
    with np.load(os.path.join(GOLD, name)) as f:
        return {k: torch.from_numpy(f[k]) for k in f.files}


def full_state(configs, seed, dense):
    shapes = {}
    for slot, mlp_cfg in orc.model_slots(configs).items():
        for k, v in orc.MlpSpec(mlp_cfg).param_shapes().items():
            shapes[f'{slot}.{k}'] = v
    state = orc.deterministic_state(shapes, seed)
    return synthetic.densify_state(state) if dense else state


def checksum(state):
    return np.array([sum(float(v.double().sum()) for v in state.values()),
                     sum(float(v.double().abs().sum()) for v in state.values())])


def render_case(name):
    """-> configs, state_dict, input batch, randoms table, golden arrays."""
    kind, ndc, camera = RENDER_CASES[name]
    g = load(name)
    seed, n, dense, _ = [int(v) for v in g['meta']]
    configs = synthetic.make_configs(kind, ndc=ndc)
    state = full_state(configs, seed, bool(dense))
    np.testing.assert_allclose(checksum(state), g['checksum'].numpy(), rtol=1e-12)
    batch = {k[3:]: v for k, v in g.items() if k.startswith('in_')}
    batch['iter_num'], batch['num_frames'] = 0, 3
    table = {k[4:]: v for k, v in g.items() if k.startswith('rnd_')}
    return configs, state, batch, table, g
