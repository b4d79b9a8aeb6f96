"""TEST INFRASTRUCTURE — not product code.

Imports the *unmodified* reference ``deepspeech_pytorch/model.py`` from
``/root/reference`` in a container that lacks pytorch_lightning / hydra /
omegaconf / torchmetrics / Levenshtein, by satisfying those imports with inert
stand-ins.  Only the imports are stubbed; the
arithmetic that runs is the reference's own ``nn.Module`` code on torch CPU.

This module is used ONLY by ``oracle/make_golden.py`` (dev container, where
``/root/reference`` exists) to generate the fixtures under ``tests/golden/`` and
to pin ``oracle/ds2_oracle.py``.  Nothing on the GPU box imports it.
"""
import sys
import types
from dataclasses import dataclass

import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


def _install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):  # reference model.py:138,147,270 use these hooks only
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        sys.modules["pytorch_lightning"] = pl
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class OmegaConf:
            get_type = staticmethod(type)  # reference model.py:152,274,282

        oc.OmegaConf, oc.MISSING = OmegaConf, "???"
        sys.modules["omegaconf"] = oc
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class Metric(nn.Module):  # reference validation.py:13,63
            def __init__(self, dist_sync_on_step=False):
                super().__init__()

            def add_state(self, name, default, dist_reduce_fx=None):
                setattr(self, name, default)

        tm.Metric = Metric
        sys.modules["torchmetrics"] = tm
    if "Levenshtein" not in sys.modules:
        lev = types.ModuleType("Levenshtein")
        lev.distance = lambda a, b: 0
        sys.modules["Levenshtein"] = lev


def load_reference():
    """Returns a namespace with the reference's model classes and config dataclasses."""
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from deepspeech_pytorch.enums import SpectrogramWindow, RNNType  # imports cleanly

    name = "deepspeech_pytorch.configs.train_config"
    if name not in sys.modules:
        # reference configs/train_config.py:41 is illegal on python>=3.11 (mutable dataclass
        # default); same fields/defaults restated here for the import only.
        tc = types.ModuleType(name)

        @dataclass
        class SpectConfig:
            sample_rate: int = 16000
            window_size: float = .02
            window_stride: float = .01
            window: SpectrogramWindow = SpectrogramWindow.hamming

        @dataclass
        class BiDirectionalConfig:
            rnn_type: RNNType = RNNType.lstm
            hidden_size: int = 1024
            hidden_layers: int = 5

        @dataclass
        class UniDirectionalConfig(BiDirectionalConfig):
            lookahead_context: int = 20

        @dataclass
        class OptimConfig:
            learning_rate: float = 1.5e-4
            learning_anneal: float = 0.99
            weight_decay: float = 1e-5

        @dataclass
        class SGDConfig(OptimConfig):
            momentum: float = 0.9

        @dataclass
        class AdamConfig(OptimConfig):
            eps: float = 1e-8
            betas: tuple = (0.9, 0.999)

        for c in (SpectConfig, BiDirectionalConfig, UniDirectionalConfig, OptimConfig, SGDConfig, AdamConfig):
            setattr(tc, c.__name__, c)
        sys.modules[name] = tc
    tc = sys.modules[name]
    from deepspeech_pytorch import model as ref_model
    from deepspeech_pytorch.decoder import GreedyDecoder

    ns = types.SimpleNamespace(
        model=ref_model, DeepSpeech=ref_model.DeepSpeech, MaskConv=ref_model.MaskConv,
        BatchRNN=ref_model.BatchRNN, Lookahead=ref_model.Lookahead, SequenceWise=ref_model.SequenceWise,
        GreedyDecoder=GreedyDecoder, RNNType=RNNType, SpectConfig=tc.SpectConfig,
        BiDirectionalConfig=tc.BiDirectionalConfig, UniDirectionalConfig=tc.UniDirectionalConfig,
        SGDConfig=tc.SGDConfig, AdamConfig=tc.AdamConfig)
    return ns
