"""Word / character error rate accumulators of the validation loop — the `self.wer` / `self.cer` attributes of the
reference model (deepspeech_pytorch/model.py:204-212,251-271) with the semantics of
deepspeech_pytorch/validation.py:13-126: decode predictions with the evaluation decoder (here the GPU greedy
decoder), turn the flat int64 targets back into strings, accumulate edit distances and reference lengths;
`compute()` returns the rate in percent.

torchmetrics and python-Levenshtein are absent from this image: the accumulators are plain `nn.Module`s with
tensor states (so they follow `.to(device)` and appear under Lightning like the reference's metrics) and the edit
distance is a two-row dynamic programme over integer sequences.
"""
import torch
import torch.nn as nn


def edit_distance(a, b) -> int:
    """Levenshtein distance between two sequences (unit costs) — Lev.distance of validation.py:80,122"""
    if len(a) < len(b):
        a, b = b, a
    if not b:
        return len(a)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


class ErrorRate(nn.Module):
    """validation.py:13-45"""

    def __init__(self, decoder, target_decoder, save_output: bool = False):
        super().__init__()
        self.decoder, self.target_decoder, self.save_output = decoder, target_decoder, save_output

    def calculate_metric(self, transcript, reference):
        raise NotImplementedError

    @torch.no_grad()
    def update(self, preds, preds_sizes, targets, target_sizes):
        split_targets, offset = [], 0
        targets = torch.as_tensor(targets).cpu()
        for size in torch.as_tensor(target_sizes).tolist():
            split_targets.append(targets[offset:offset + size])
            offset += size
        decoded_output, _ = self.decoder.decode(preds, preds_sizes)
        target_strings = self.target_decoder.convert_to_strings(split_targets)
        for x in range(len(target_strings)):
            self.calculate_metric(transcript=decoded_output[x][0], reference=target_strings[x][0])

    def forward(self, preds, preds_sizes, targets, target_sizes):
        self.update(preds, preds_sizes, targets, target_sizes)
        return self.compute()


class CharErrorRate(ErrorRate):
    """validation.py:48-81: edit distance over characters with spaces removed / number of reference characters"""

    def __init__(self, decoder, target_decoder, save_output: bool = False):
        super().__init__(decoder, target_decoder, save_output)
        self.register_buffer("cer", torch.tensor(0), persistent=False)
        self.register_buffer("n_chars", torch.tensor(0), persistent=False)

    def cer_calc(self, s1, s2):
        return edit_distance(s1.replace(' ', ''), s2.replace(' ', ''))

    def calculate_metric(self, transcript, reference):
        self.cer += self.cer_calc(transcript, reference)
        self.n_chars += len(reference.replace(' ', ''))

    def compute(self):
        return float(self.cer) / max(1, int(self.n_chars)) * 100

    def reset(self):
        self.cer.zero_()
        self.n_chars.zero_()


class WordErrorRate(ErrorRate):
    """validation.py:84-126: edit distance over whitespace-separated words / number of reference words"""

    def __init__(self, decoder, target_decoder, save_output: bool = False):
        super().__init__(decoder, target_decoder, save_output)
        self.register_buffer("wer", torch.tensor(0), persistent=False)
        self.register_buffer("n_tokens", torch.tensor(0), persistent=False)

    def wer_calc(self, s1, s2):
        return edit_distance(s1.split(), s2.split())

    def calculate_metric(self, transcript, reference):
        self.wer += self.wer_calc(transcript, reference)
        self.n_tokens += len(reference.split())

    def compute(self):
        return float(self.wer) / max(1, int(self.n_tokens)) * 100

    def reset(self):
        self.wer.zero_()
        self.n_tokens.zero_()
