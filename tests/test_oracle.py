"""The oracle (oracle/ds2_oracle.py) against outputs of the reference itself (tests/golden, written by
oracle/make_golden.py from the unmodified reference model.py).  CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from oracle import ds2_oracle as O

TOL = 2e-5  # fp32 CPU vs fp32 CPU, different summation orders only


def test_seq_lens_integer_exact():
    # reference model.py:299-310 probed in SURVEY Appendix B
    L = torch.tensor([1, 2, 3, 100, 999, 1000, 1001, 4000])
    assert O.get_seq_lens(L).tolist() == [1, 1, 2, 50, 500, 500, 501, 2000]
    assert O.get_seq_lens(L).dtype == torch.int32


def test_input_sizes_truncation_quirk():
    # reference model.py:243: f32 multiply + trunc loses a frame for 5 lengths at T=1000
    T = 1000
    lens = torch.arange(1, T + 1)
    pct = (lens.double() / T).float()
    got = O.input_sizes_from_percentages(pct, T)
    bad = lens[got.long() != lens].tolist()
    assert bad == [251, 253, 502, 506, 511]


def test_forward_train_matches_reference(golden):
    cfg = golden.oracle_cfg()
    taps = {}
    out, out_lens, _, nb = O.forward(golden.x, golden.input_sizes, golden.params, cfg, training=True, taps=taps)
    assert out_lens.tolist() == golden.output_sizes.tolist()
    assert out.shape == golden.logits.shape
    assert rel_err(out, golden.logits) < TOL
    assert rel_err(taps["conv"], golden.taps["conv"]) < TOL
    for i in range(cfg.hidden_layers):
        assert rel_err(taps[f"rnn{i}"], golden.taps[f"rnn{i}"]) < TOL
    for k, v in nb.items():
        assert rel_err(v, golden.buffers_after[k]) < TOL, k


def test_train_step_matches_reference(golden):
    cfg = golden.oracle_cfg()
    r = O.train_step(golden.x, golden.targets, golden.input_percentages.clone(), golden.target_sizes,
                     golden.params, cfg)
    assert abs(r["loss"] - golden.loss) <= 2e-5 * max(1.0, abs(golden.loss))
    assert set(r["grads"].keys()) == set(golden.grads.keys())
    for k, g in golden.grads.items():
        assert rel_err(r["grads"][k], g) < 5e-5, k


def test_aten_rnn_path_equals_explicit_cells(golden):
    cfg = golden.oracle_cfg()
    a = O.train_step(golden.x, golden.targets, golden.input_percentages.clone(), golden.target_sizes,
                     golden.params, cfg, use_aten_rnn=True, use_aten_ctc=True)
    assert abs(a["loss"] - golden.loss) <= 2e-5 * max(1.0, abs(golden.loss))
    for k, g in golden.grads.items():
        assert rel_err(a["grads"][k], g) < 5e-5, k


def test_eval_forward_and_states(golden):
    cfg = golden.oracle_cfg()
    P = dict(golden.params)
    P.update(golden.buffers_after)  # the eval forward in make_golden ran after the train step
    out, _, hs, _ = O.forward(golden.x, golden.input_sizes, P, cfg, training=False)
    assert rel_err(out, golden.eval_out) < TOL
    for i, h in enumerate(hs):
        hn = h[0] if isinstance(h, tuple) else h
        assert rel_err(hn, torch.from_numpy(golden.z[f"eval_hn/{i}"])) < TOL
        if isinstance(h, tuple):
            assert rel_err(h[1], torch.from_numpy(golden.z[f"eval_cn/{i}"])) < TOL
    # integer outputs: argmax path and greedy offsets are bit-exact (reference decoder.py:144-181)
    assert out.argmax(2).tolist() == golden.z["eval_argmax"].tolist()
    paths = O.greedy_path(out, golden.output_sizes)
    for b, (_, offs) in enumerate(paths):
        assert offs == golden.z[f"eval_offsets/{b}"].tolist()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ctc_numpy_vs_aten(seed):
    g = torch.Generator().manual_seed(seed)
    T, B, C = 40, 5, 29
    logits = torch.randn(T, B, C, generator=g, dtype=torch.float64, requires_grad=True)
    in_len = torch.tensor([40, 33, 17, 9, 5], dtype=torch.int32)
    tg_len = torch.tensor([12, 1, 9, 4, 5], dtype=torch.int32)   # utt 2: 9 labels in 17 frames, utt 4: L==T
    targets = torch.randint(1, C, (int(tg_len.sum()),), generator=g)
    targets[13:22] = torch.tensor([4, 4, 4, 4, 4, 4, 4, 4, 4])   # 9 repeats need 17 frames: exactly feasible
    targets[26:31] = torch.tensor([7, 7, 1, 2, 3])                # L=5 with a repeat in T=5: infeasible
    loss = F.ctc_loss(logits.log_softmax(-1), targets, in_len, tg_len, blank=0, reduction="none",
                      zero_infinity=True)
    loss.sum().backward()
    nll, grad = O.ctc_loss_and_grad(logits.detach().numpy(), targets.numpy(), in_len.numpy(), tg_len.numpy())
    assert np.allclose(nll, loss.detach().numpy(), rtol=1e-10, atol=1e-10)
    assert nll[4] == 0.0 and np.all(grad[:, 4] == 0.0)
    assert np.allclose(grad, logits.grad.numpy(), rtol=1e-9, atol=1e-10)
    for b in range(B):
        assert np.all(grad[int(in_len[b]):, b] == 0.0)


def test_lookahead_indexing():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(9, 2, 4, generator=g)
    w = torch.randn(4, 1, 3, generator=g)
    y = O.lookahead(x, w)
    for t in range(9):
        acc = torch.zeros(2, 4)
        for k in range(3):
            if t + k < 9:
                acc += w[:, 0, k] * x[t + k]
        assert torch.allclose(y[t], acc, atol=1e-6)
