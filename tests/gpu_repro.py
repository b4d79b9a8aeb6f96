import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import deepspeech_pytorch_b200 as ds
from gpu_helpers import make_model, rel
ds.set_precision("tf32")
m = make_model("lstm", True, 32, 1).train()
sm = m.conv.seq_module
B, T = 3, 248
x = torch.randn(B, 1, 161, T).cuda()
ol = m.get_seq_lens(torch.tensor([T, T - 9, T - 40])).cuda()
args = [sm[0].weight, sm[0].bias, sm[1].weight, sm[1].bias, sm[1].running_mean, sm[1].running_var, sm[3].weight,
        sm[3].bias, sm[4].weight, sm[4].bias, sm[4].running_mean, sm[4].running_var]
def run(env):
    for k in ("DS2_NO_CONV_TC", "DS2_NO_CONV_TC_FWD", "DS2_NO_CONV_TC_DGRAD", "DS2_NO_CONV_TC_WGRAD"):
        os.environ.pop(k, None)
    for k in env:
        os.environ[k] = "1"
    for p in m.parameters():
        p.grad = None
    y = ds.ops.ConvFrontend.apply(x, ol, *args, True, 0.1, 1e-5)
    g = torch.Generator(device="cuda").manual_seed(1)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    y.backward(dy)
    torch.cuda.synchronize()
    return (y.detach().clone(), [sm[i].weight.grad.clone() for i in (0, 1, 3, 4)], [sm[i].bias.grad.clone() for i in (0, 1, 3, 4)])
ref = run(["DS2_NO_CONV_TC"])
for name, env in (("fwd only TC", ["DS2_NO_CONV_TC_DGRAD", "DS2_NO_CONV_TC_WGRAD"]),
                  ("dgrad only TC", ["DS2_NO_CONV_TC_FWD", "DS2_NO_CONV_TC_WGRAD"]),
                  ("wgrad only TC", ["DS2_NO_CONV_TC_FWD", "DS2_NO_CONV_TC_DGRAD"])):
    r = run(env)
    print(name, "y", f"{rel(r[0], ref[0]):.2e}", "dW", [f"{rel(a, b):.2e}" for a, b in zip(r[1], ref[1])],
          "db", [f"{rel(a, b):.2e}" for a, b in zip(r[2], ref[2])])
