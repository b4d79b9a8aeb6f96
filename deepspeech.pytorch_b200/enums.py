"""Same members as the reference's deepspeech_pytorch/enums.py:6-21 (RNNType values are the torch
classes, because the reference's configs and checkpoints carry them that way)."""
from enum import Enum

from torch import nn


class DecoderType(Enum):
    greedy = 'greedy'
    beam = 'beam'


class SpectrogramWindow(Enum):
    hamming = 'hamming'
    hann = 'hann'
    blackman = 'blackman'
    bartlett = 'bartlett'


class RNNType(Enum):
    lstm = nn.LSTM
    rnn = nn.RNN
    gru = nn.GRU
