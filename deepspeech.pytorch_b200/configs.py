"""Structured configs with the field names and defaults of the reference's
deepspeech_pytorch/configs/train_config.py:16-73, restated so that they are legal on
python >= 3.11 (the reference's mutable dataclass defaults at :41,:87-89 are not) and usable
without hydra/omegaconf installed."""
from dataclasses import dataclass, field

from .enums import RNNType, SpectrogramWindow


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


@dataclass
class AugmentationConfig:
    speed_volume_perturb: bool = False
    spec_augment: bool = False
    noise_dir: str = ''
    noise_prob: float = 0.4
    noise_min: float = 0.0
    noise_max: float = 0.5


@dataclass
class DataConfig:
    train_path: str = 'data/train_manifest.csv'
    val_path: str = 'data/val_manifest.csv'
    batch_size: int = 64
    num_workers: int = 4
    labels_path: str = 'labels.json'
    spect: SpectConfig = field(default_factory=SpectConfig)
    augmentation: AugmentationConfig = field(default_factory=AugmentationConfig)
    prepare_data_per_node: bool = True


def cfg_type(cfg):
    """OmegaConf.get_type(cfg) when omegaconf wraps the config, else type(cfg) (model.py:152,274,282)."""
    try:
        from omegaconf import OmegaConf  # optional
        t = OmegaConf.get_type(cfg)
        if t is not None:
            return t
    except Exception:
        pass
    return type(cfg)


def is_kind(cfg, *names):
    """dataclass-type dispatch by class name, so configs built from the reference's own
    deepspeech_pytorch.configs.train_config classes are accepted as well."""
    return any(c.__name__ in names for c in cfg_type(cfg).__mro__)
