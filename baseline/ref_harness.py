"""Drives the UNMODIFIED reference (NagabhushanSN95/SimpleNeRF, `src/`) on synthetic data: test and measurement
infrastructure only -- nothing under simplenerf_b200/ imports this module.

Where the reference comes from: `baseline/_ref/src` (a verbatim copy that `__graft_entry__.build()` makes from
/root/reference/src when that exists; git-ignored, it travels to the GPU box with the snapshot) or, in the build
container, /root/reference/src itself.  `available()` says whether either exists; callers skip / report otherwise.

What is stubbed: five imports the reference's modules pull in but the training step never calls when
`downsampling_factor == 1` (test 4): skimage(.io, .transform), simplejson (-> json), deepdiff.DeepDiff,
matplotlib(.pyplot), skvideo(.io); and torch.utils.tensorboard.SummaryWriter when tensorboard is not installed.
What is synthetic: the `raw_data_dict` the reference's data loader would read from disk (3 random images, cameras on a
line, random sparse-depth tables) -- everything downstream of it (pose recentring, ray cache, batching, model, losses, Adam,
`Trainer.train_one_iter`, src/Trainer01.py:61-107) is the reference's own code.
"""
from __future__ import annotations

import copy
import importlib
import json
import os
import sys
import tempfile
import types
from pathlib import Path
from typing import Optional

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
CANDIDATES = (ROOT / 'baseline' / '_ref', Path('/root/reference'))


def ref_root() -> Optional[Path]:
    for c in CANDIDATES:
        if (c / 'src' / 'models' / 'SimpleNeRF01.py').exists():
            return c
    return None


def available() -> bool:
    return ref_root() is not None


def install_stubs() -> None:
    def mod(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        try:
            return importlib.import_module(name)
        except Exception:   # noqa: BLE001  (absent, or present but broken in this image)
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
            return m

    def unavailable(*_a, **_k):
        raise RuntimeError('stubbed optional dependency of the reference was called (it is not on the training-step path)')

    sk = mod('skimage')
    sk.io = mod('skimage.io', imread=unavailable, imsave=unavailable)
    sk.transform = mod('skimage.transform', rescale=unavailable, resize=unavailable)
    mod('simplejson', dump=json.dump, dumps=json.dumps, load=json.load, loads=json.loads)
    mod('deepdiff', DeepDiff=lambda a, b, **k: {} if a == b else {'changed': True})
    mp = mod('matplotlib')
    mp.pyplot = mod('matplotlib.pyplot')
    sv = mod('skvideo')
    sv.io = mod('skvideo.io')
    try:
        from torch.utils.tensorboard import SummaryWriter  # noqa: F401
    except Exception:   # noqa: BLE001
        tb = types.ModuleType('torch.utils.tensorboard')

        class SummaryWriter:      # the step under test never logs
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

            def close(self):
                pass
        tb.SummaryWriter = SummaryWriter
        sys.modules['torch.utils.tensorboard'] = tb


def use_reference() -> Path:
    root = ref_root()
    if root is None:
        raise RuntimeError('the reference tree is not available (neither baseline/_ref nor /root/reference)')
    install_stubs()
    src = str(root / 'src')
    if src not in sys.path:
        sys.path.insert(0, src)
    return root


def register_dropin() -> None:
    """What a maintainer does by dropping `models/FusedSimpleNeRF01.py` into the reference tree: make
    `models.FusedSimpleNeRF01` importable so that `ModelFactory.get_model` finds it by name (src/models/ModelFactory.py:10-22)."""
    use_reference()
    importlib.import_module('models')
    sys.modules['models.FusedSimpleNeRF01'] = importlib.import_module('simplenerf_b200.models.FusedSimpleNeRF01')


def load_configs(train_num: int = 1021) -> dict:
    root = use_reference()
    with open(root / 'runs' / 'training' / f'train{train_num:04}' / 'Configs.json') as f:
        return json.load(f)


def make_raw_data(resolution=(48, 64), n_views: int = 3, seed: int = 0, sparse_points: int = 200) -> dict:
    """The dict `NerfLlffDataLoader.load_data` returns (src/data_loaders/NerfLlffDataLoader01.py:30), filled with synthetic content."""
    import pandas
    h, w = resolution
    rng = np.random.RandomState(seed)
    images = rng.randint(0, 256, size=(n_views, h, w, 3)).astype(np.uint8)
    extrinsics = np.stack([np.eye(4) for _ in range(n_views)]).astype(np.float64)
    extrinsics[:, 0, 3] = np.linspace(-0.5, 0.5, n_views)
    focal = 815.1316 * w / 1008.0
    intrinsic = np.array([[focal, 0, w / 2], [0, focal, h / 2], [0, 0, 1]], dtype=np.float64)
    intrinsics = np.stack([intrinsic] * n_views)
    bounds = np.array([1.3, 8.0])                                    # [min, max] over the views (NerfLlffDataLoader01.py load_nerf_data)
    sparse = {}
    for f in range(n_views):
        sparse[f] = pandas.DataFrame({'x': rng.uniform(0, w - 1, sparse_points), 'y': rng.uniform(0, h - 1, sparse_points),
                                      'depth': rng.uniform(2.0, 6.0, sparse_points),
                                      'reprojection_error': rng.uniform(0.1, 1.0, sparse_points)})
    return {'frame_nums': np.arange(n_views), 'nerf_data': {'images': images, 'extrinsics': extrinsics, 'intrinsics': intrinsics,
                                                           'bounds': bounds, 'resolution': (h, w)},
            'sparse_depth_data': sparse}


def build_trainer(model_name: str = 'SimpleNeRF01', device=(0,), resolution=(48, 64), num_rays: int = 2048,
                  sparse_rays: int = 2048, sub_batch_size: Optional[int] = 2048, seed: int = 230, model_extra: Optional[dict] = None,
                  optimizer_factory=None, loss_computer_factory=None, iter_num_for_losses: Optional[int] = None):
    """The body of `Trainer01.start_training` (src/Trainer01.py:488-527) with the disk loader replaced by `make_raw_data`.
    device: [k] like the reference's configs; the reference falls back to the CPU by itself when CUDA is not visible
    (src/utils/CommonUtils01.py:15-27) -- hide the GPUs (CUDA_VISIBLE_DEVICES='') to time its CPU path on a GPU box.  -> (trainer, configs)"""
    import torch
    use_reference()
    if model_name.startswith('Fused'):
        register_dropin()
    Trainer01 = importlib.import_module('Trainer01')
    from data_preprocessors.DataPreprocessorFactory import get_data_preprocessor
    from loss_functions.LossComputer01 import LossComputer
    from lr_decayers.LearningRateDecayerFactory import get_lr_decayer
    from models.ModelFactory import get_model

    configs = copy.deepcopy(load_configs())
    configs['device'] = list(device)
    configs['model']['name'] = model_name
    if model_extra:
        configs['model'].update(model_extra)
    configs['data_loader']['num_rays'] = num_rays
    configs['data_loader']['sparse_depth']['num_rays'] = sparse_rays
    if sub_batch_size is None:
        configs.pop('sub_batch_size', None)
    else:
        configs['sub_batch_size'] = sub_batch_size
    configs['seed'] = seed
    out_dir = Path(tempfile.mkdtemp(prefix='snerf_ref_'))
    configs['root_dirpath'] = out_dir
    configs['output_dirpath'] = out_dir
    Trainer01.init_seeds(seed)                                                    # :477-485
    pre = get_data_preprocessor(configs, mode='train', raw_data_dict=make_raw_data(resolution, seed=seed))
    model_configs = pre.get_model_configs()
    model = get_model(configs, model_configs)
    model = torch.nn.DataParallel(model, device_ids=configs['device'])            # :514
    loss_computer = (loss_computer_factory or LossComputer)(configs)
    make_opt = optimizer_factory or (lambda params: torch.optim.Adam(params, lr=configs['optimizer']['lr_initial'],
                                                                     betas=(configs['optimizer']['beta1'], configs['optimizer']['beta2'])))
    optimizer = make_opt(list(model.parameters()))                               # :516
    trainer = Trainer01.Trainer(configs, model_configs, pre, None, model, loss_computer, optimizer, get_lr_decayer(configs),
                                out_dir, configs['device'])
    return trainer, configs
