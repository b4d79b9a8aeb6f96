"""CPU: host-side logic of the shell that needs no GPU — WER/CER accumulators (reference validation.py:13-126),
target-string conversion (decoder.py:125-163), the bench's synthetic batch format (data_loader.py:247-270)."""
import os
import sys

import torch

from conftest import ROOT

import deepspeech_pytorch_b200 as ds
from deepspeech_pytorch_b200.metrics import CharErrorRate, WordErrorRate, edit_distance


def test_edit_distance_known_answers():
    assert edit_distance("kitten", "sitting") == 3
    assert edit_distance("", "abc") == 3 and edit_distance("abc", "") == 3 and edit_distance("", "") == 0
    assert edit_distance("flaw", "lawn") == 2
    assert edit_distance("the cat sat".split(), "the cat sat".split()) == 0
    assert edit_distance("the cat sat".split(), "the bat sat down".split()) == 2


def test_convert_to_strings_drops_blank_and_optionally_collapses():
    dec = ds.GreedyDecoder(ds.LABELS)
    idx = {c: i for i, c in enumerate(ds.LABELS)}
    seq = torch.tensor([idx["H"], idx["H"], 0, idx["I"], idx[" "], idx["Y"], idx["O"], idx["O"]])
    assert dec.convert_to_strings([seq]) == [["HHI YOO"]]
    strings, offs = dec.convert_to_strings([seq], remove_repetitions=True, return_offsets=True)
    assert strings == [["HI YO"]] and offs[0][0].tolist() == [0, 3, 4, 5, 6]
    assert dec.convert_to_strings([seq], sizes=[3]) == [["HH"]]


class _FakeDecoder:
    """stands in for the GPU greedy decoder: returns fixed hypotheses"""

    def __init__(self, hyps, real):
        self.hyps, self.real = hyps, real

    def decode(self, preds, sizes):
        return [[h] for h in self.hyps], None

    def convert_to_strings(self, seqs):
        return self.real.convert_to_strings(seqs)


def test_wer_cer_accumulate_like_the_reference_metrics():
    real = ds.GreedyDecoder(ds.LABELS)
    idx = {c: i for i, c in enumerate(ds.LABELS)}
    refs = ["THE CAT SAT", "A DOG"]
    hyps = ["THE BAT SAT DOWN", "A DOG"]
    targets = torch.tensor([idx[c] for r in refs for c in r])
    sizes = torch.tensor([len(r) for r in refs], dtype=torch.int32)
    dec = _FakeDecoder(hyps, real)
    wer, cer = WordErrorRate(dec, dec), CharErrorRate(dec, dec)
    w = wer(None, None, targets, sizes)
    c = cer(None, None, targets, sizes)
    # words: 2 errors / 5 reference words; characters (spaces removed): C->B and +DOWN = 5 edits / 13 characters
    assert abs(w - 100 * 2 / 5) < 1e-9 and abs(c - 100 * 5 / 13) < 1e-9
    assert int(wer.n_tokens) == 5 and int(cer.n_chars) == 13
    wer(None, None, targets, sizes)                         # accumulates over batches like a torchmetrics Metric
    assert int(wer.wer) == 4 and int(wer.n_tokens) == 10 and abs(wer.compute() - 40.0) < 1e-9
    wer.reset(); cer.reset()
    assert wer.compute() == 0.0 and cer.compute() == 0.0


def test_model_exposes_wer_cer_without_changing_the_state_dict():
    m = ds.DeepSpeech(ds.LABELS, ds.BiDirectionalConfig(hidden_size=8, hidden_layers=1), 32, ds.AdamConfig(),
                      ds.SpectConfig())
    assert isinstance(m.wer, WordErrorRate) and isinstance(m.cer, CharErrorRate)
    assert not any("wer" in k or "cer" in k for k in m.state_dict())


def test_bench_synthetic_batch_has_the_collate_format():
    sys.path.insert(0, ROOT)
    import bench
    x, targets, pct, tsz = bench.synth_batch(3, 50, 7, seed=1)
    assert x.shape == (3, 1, 161, 50) and x.dtype == torch.float32
    assert targets.dtype == torch.int64 and targets.shape == (21,) and int(targets.min()) >= 1 and int(targets.max()) <= 28
    assert pct.dtype == torch.float32 and pct.tolist() == [1.0, 1.0, 1.0]
    assert tsz.dtype == torch.int32 and tsz.tolist() == [7, 7, 7]
    assert "oracle" not in open(os.path.join(ROOT, "bench.py")).read().split("def run_b200")[1].split("def main")[0] \
        .replace("oracle port", "").replace("_oracle_cfg", "")
