"""Pin oracle/train_oracle.py against the UNMODIFIED reference model in training mode and write
tests/golden/i3d_train.npz (authoring container only: needs /root/reference; CPU, fp32).

The reference training step (pt/train_i3d_smth.py:192-250) is model.train(); output = model(input);
loss = CrossEntropyLoss(output, target); loss.backward(); optimizer.step().  Here: the reference's own
I3D_doubled.Model (dropout disabled through its own constructor argument, soft_max 0 as pt/configs/config_i3d_smth.py
sets it) on two seeded 16x224x224 clips, torch.optim.SGD(lr 0.01, momentum 0.9, weight_decay 1e-5), two steps.
A gradient tensor is stored as its L2 norm plus 64 entries at seeded positions (the full set is 49 MB).
    python oracle/pin_train_step.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/video_features_pytorch"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from oracle import synthetic, train_oracle  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
LR, MOM, WD = 0.01, 0.9, 1e-5


sample_index = train_oracle.sample_index


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    from models import I3D_doubled
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = I3D_doubled.Model(174, last_stride=1, stride_mod_layers="", softMax=0, dropout_keep_prob=0.0)
    sd0 = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    x = synthetic.clips(2)  # [2,3,16,224,224], 0..255
    target = torch.tensor([3, 100])
    ref.train()
    opt = torch.optim.SGD(ref.parameters(), lr=LR, momentum=MOM, weight_decay=WD)
    crit = torch.nn.CrossEntropyLoss()
    gold = {"target": target.numpy(), "lr": LR, "momentum": MOM, "weight_decay": WD}
    sd, state = dict(sd0), None
    for step in (1, 2):
        ref.zero_grad()
        out = ref(x)
        loss = crit(out, target)
        opt.zero_grad()
        loss.backward()
        g_ref = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
        # the oracle on the same state
        l_or, logits_or, g_or, buf_or = train_oracle.loss_and_grads(sd, x, target)
        print("step %d: loss ref %.6f oracle %.6f" % (step, loss.item(), l_or))
        assert abs(l_or - loss.item()) < 1e-5 * max(1.0, abs(loss.item()))
        assert rel(logits_or, out.detach()) < 1e-5
        worst = max((rel(g_or[k], g_ref[k]), k) for k in g_ref)
        print("  worst gradient mismatch oracle vs reference: %.2e (%s)" % worst)
        assert worst[0] < 2e-4, worst
        # fp32 on a deep BatchNorm network is itself noisy (the gradient of a BatchNorm bias that feeds another
        # convolution + BatchNorm is a sum that cancels almost completely): the SAME state evaluated in fp64 shows how
        # far the reference's fp32 numbers are from the exact ones, tensor by tensor - the tests hold the kernels to
        # the fp64 values within a multiple of that distance
        _, _, g_64, _ = train_oracle.loss_and_grads(sd, x, target, dtype=torch.float64)
        unc = sorted(((rel(g_ref[k], g_64[k]), k) for k in g_ref), reverse=True)
        print("  reference fp32 vs fp64 oracle: worst %.2e (%s), median %.2e" % (unc[0] + (unc[len(unc) // 2][0],)))
        for k, g in g_64.items():
            gold["g64norm_%d/%s" % (step, k)] = np.float64(g.norm().item())
            gold["g64samp_%d/%s" % (step, k)] = g.flatten()[sample_index(k, g.numel())].numpy()
        opt.step()
        new_ref = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        new_or, state = train_oracle.sgd_step(sd, g_or, LR, MOM, WD, state)
        worst = max((rel(new_or[k], new_ref[k]), k) for k in new_or)
        print("  worst parameter mismatch after the update: %.2e (%s)" % worst)
        assert worst[0] < 1e-5, worst
        for k, v in buf_or.items():
            assert rel(v, new_ref[k]) < 1e-5, k
        gold["loss_%d" % step] = np.float32(loss.item())
        gold["logits_%d" % step] = out.detach().numpy()
        for k, g in g_ref.items():
            gold["gnorm_%d/%s" % (step, k)] = np.float32(g.norm().item())
            gold["gsamp_%d/%s" % (step, k)] = g.flatten()[sample_index(k, g.numel())].numpy()
        for k, v in new_ref.items():
            if train_oracle.is_param(k) or ".bn.running_" in k:
                gold["psamp_%d/%s" % (step, k)] = v.flatten()[sample_index(k, v.numel())].numpy()
        sd = new_ref  # the second step starts from the reference's updated state
    np.savez_compressed(os.path.join(GOLD, "i3d_train.npz"), **gold)
    print("train-step oracle pinned against the reference; tests/golden/i3d_train.npz written")


if __name__ == "__main__":
    main()
