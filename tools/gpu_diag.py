"""One-shot GPU diagnostics: runs every kernel family against the oracle and prints a detailed report.

Used during bring-up so that ONE gpurun call tells which piece is wrong and how (error structure, not just pass/fail).
Usage: python tools/gpu_diag.py [section ...]   sections: conv softargmax potentials forward
"""
import os
import sys
import time
import traceback

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deepgraphpose_b200 import engine as E, synthetic  # noqa: E402
from oracle import dgp_ops, pose_net, tf_ops  # noqa: E402


def bf16_round(a):
    return torch.from_numpy(np.asarray(a, np.float32)).to(torch.bfloat16).float()


def conv_ref(x_bf16_f32, w, stride, dil, pad_mode, scale, shift, residual, res_sub, relu):
    wq = bf16_round(w)
    if pad_mode == 0:
        y = tf_ops.conv2d(x_bf16_f32, wq, stride, dil, "SAME")
    elif pad_mode == 1:
        y = tf_ops.conv2d_same(x_bf16_f32, wq, stride, dil)
    else:
        y = tf_ops.conv2d(x_bf16_f32, wq, stride, dil, "VALID")
    if scale is not None:
        y = y * torch.from_numpy(scale)
    if shift is not None:
        y = y + torch.from_numpy(shift)
    if residual is not None:
        y = y + residual[:, ::res_sub, ::res_sub, :]
    if relu:
        y = torch.relu(y)
    return y


def report_err(name, got, ref):
    got = got.float().cpu()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-9
    print("  %-44s max_abs_err %.4g (ref max %.4g) rel %.3g  nan %d" % (
        name, err.max().item(), denom, err.max().item() / denom, int(torch.isnan(got).sum())), flush=True)
    return err.max().item() / denom


def diag_conv(eng, only=None):
    print("== conv (tcgen05 implicit GEMM) ==", flush=True)
    rng = np.random.default_rng(0)
    cases = [
        # name, N,H,W,Cin,Cout,R,stride,dil,pad_mode,bn,res,res_sub,relu,block_n,out_f32
        ("1x1 64->64 M=128 exact", 1, 8, 16, 64, 64, 1, 1, 1, 0, False, False, 1, False, 0, True),
        ("1x1 64->64 tail", 2, 13, 17, 64, 64, 1, 1, 1, 0, False, False, 1, False, 0, True),
        ("1x1 256->128 bn relu", 2, 13, 17, 256, 128, 1, 1, 1, 0, True, False, 1, True, 0, False),
        ("1x1 128->512 bn res relu", 2, 13, 17, 128, 512, 1, 1, 1, 0, True, True, 1, True, 0, False),
        ("1x1 128->512 block_n 128", 2, 13, 17, 128, 512, 1, 1, 1, 0, True, True, 1, True, 128, False),
        ("1x1 512->48 block_n 48", 1, 9, 11, 512, 48, 1, 1, 1, 0, False, False, 1, False, 48, True),
        ("1x1 K=2048 ->256 many kblocks", 1, 20, 20, 2048, 256, 1, 1, 1, 0, True, False, 1, True, 0, False),
        ("3x3 s1 64->64 SAME", 2, 13, 17, 64, 64, 3, 1, 1, 1, True, False, 1, True, 0, False),
        ("3x3 s1 128->128 SAME f32", 1, 21, 19, 128, 128, 3, 1, 1, 1, False, False, 1, False, 0, True),
        ("3x3 s2 64->64 conv2d_same odd", 2, 13, 17, 64, 64, 3, 2, 1, 1, True, False, 1, True, 0, False),
        ("3x3 s2 128->128 conv2d_same even", 1, 20, 18, 128, 128, 3, 2, 1, 1, True, False, 1, True, 0, False),
        ("3x3 dil2 512->512", 1, 15, 19, 512, 512, 3, 1, 2, 1, True, False, 1, True, 0, False),
        ("1x1 res subsample 2", 2, 7, 9, 64, 256, 1, 1, 1, 0, True, True, 2, True, 0, False),
        ("big 1x1 256->1024 M~20k", 4, 70, 72, 256, 1024, 1, 1, 1, 0, True, True, 1, True, 0, False),
        ("big 3x3 256->256 M~20k", 4, 70, 72, 256, 256, 3, 1, 1, 1, True, False, 1, True, 0, False),
    ]
    worst = 0
    if only is not None:
        cases = [cases[only]]
    for (name, N, H, W, Cin, Cout, R, stride, dil, pm, bn, res, res_sub, relu, block_n, out_f32) in cases:
        try:
            x = torch.from_numpy(rng.standard_normal((N, H, W, Cin)).astype(np.float32)).to(torch.bfloat16)
            w = (rng.standard_normal((R, R, Cin, Cout)) * np.sqrt(1.0 / (R * R * Cin))).astype(np.float32)
            scale = rng.uniform(0.5, 1.5, Cout).astype(np.float32) if bn else None
            shift = rng.normal(0, 0.2, Cout).astype(np.float32) if bn else None
            Ho, Wo = (-(-H // stride), -(-W // stride))
            residual = None
            if res:
                residual = torch.from_numpy(rng.standard_normal((N, Ho * res_sub, Wo * res_sub, Cout)).astype(np.float32)).to(torch.bfloat16)
            ref = conv_ref(x.float(), w, stride, dil, pm, scale, shift, residual.float() if res else None, res_sub, relu)
            got = eng.conv2d(x.cuda(), w, stride, dil, pm, scale, shift, residual.cuda() if res else None, res_sub, relu,
                             out_f32, block_n)
            torch.cuda.synchronize()
            rel = report_err(name, got, ref)
            worst = max(worst, rel)
            if rel > 2e-2:
                g = got.float().cpu()
                e = (g - ref).abs().reshape(-1, Cout)
                rows = e.max(dim=1).values
                cols = e.max(dim=0).values
                print("     bad rows: %d/%d first %s ; bad cols: %d/%d first %s" % (
                    int((rows > 1e-2).sum()), rows.numel(), (rows > 1e-2).nonzero().flatten()[:12].tolist(),
                    int((cols > 1e-2).sum()), cols.numel(), (cols > 1e-2).nonzero().flatten()[:12].tolist()))
                print("     sample got", g.reshape(-1, Cout)[0, :6].tolist(), "ref", ref.reshape(-1, Cout)[0, :6].tolist())
        except Exception:
            print("  %-44s EXCEPTION" % name)
            traceback.print_exc()
            worst = 1e9
    print("conv worst rel err %.3g" % worst, flush=True)


def diag_softargmax(eng):
    print("== softargmax / peaks ==", flush=True)
    rng = np.random.default_rng(1)
    for (B, H, W, nj, scale) in [(3, 30, 38, 4, 3.0), (2, 94, 104, 5, 4.0), (5, 60, 80, 20, 2.0), (1, 128, 160, 16, 6.0),
                                 (300, 30, 38, 4, 3.0), (2, 30, 38, 3, 30.0)]:
        try:
            logits = (rng.standard_normal((B, H, W, nj)) * scale).astype(np.float32)
            # plant clear peaks, some at borders
            for b in range(B):
                for j in range(nj):
                    r, c = rng.integers(0, H), rng.integers(0, W)
                    if j == 0:
                        r, c = 0, 0
                    if j == 1:
                        r, c = H - 1, W - 1
                    logits[b, r, c, j] += 4 * scale
            locref = rng.standard_normal((B, H, W, 2 * nj)).astype(np.float32)
            lt = torch.from_numpy(logits)
            mu_ref, _ = dgp_ops.argmax_2d_from_cm(lt, nj, 1.0, 1.0)
            out = eng.softargmax(lt.cuda(), torch.from_numpy(locref).cuda(), 1.0, 1.0)
            torch.cuda.synchronize()
            mu = out["mu"].cpu()
            dmu = (mu - mu_ref).abs().max().item()
            # peaks: oracle read-out using the GPU mu (same fp32 inputs) -> must be bit exact
            npeak_bad = 0
            lik_err = 0.0
            nb = min(B, 8)
            for b in range(nb):
                _, pk, lk = dgp_ops.estimate_pose_readout(mu[b:b + 1].numpy(), logits[b:b + 1])
                npeak_bad += int((pk != out["peak"][b].cpu().numpy()).sum())
                lik_err = max(lik_err, float(np.abs(lk - out["lik"][b].cpu().numpy()).max()))
            prob = eng.sigmoid(lt.cuda()).cpu().numpy()
            prob_ref = torch.sigmoid(lt).numpy()
            ndlc_bad = 0
            pose_err = 0.0
            for b in range(nb):
                scm, loc = pose_net.extract_cnn_output(prob[b:b + 1], locref[b:b + 1])
                pose, peaks = pose_net.argmax_pose_predict(scm, loc, 8.0)
                ndlc_bad += int((peaks != out["dlc_peak"][b].cpu().numpy()).sum())
                pose_err = max(pose_err, float(np.abs(pose - out["dlc_pose"][b].cpu().numpy()).max()))
            print("  B%d %dx%d nj%d: |mu-mu_ref| %.3g  peak mismatches %d  lik err %.3g  dlc peak mismatches %d  pose err %.3g  sigmoid err %.3g" % (
                B, H, W, nj, dmu, npeak_bad, lik_err, ndlc_bad, pose_err, float(np.abs(prob - prob_ref).max())), flush=True)
        except Exception:
            print("  softargmax case EXCEPTION", (B, H, W, nj))
            traceback.print_exc()


def diag_potentials(eng):
    print("== potentials ==", flush=True)
    rng = np.random.default_rng(2)
    T, nj = 1000, 16
    mu = torch.from_numpy(rng.uniform(0, 100, (T, nj, 2)).astype(np.float32))
    edges = synthetic.chain_skeleton(nj)
    S0 = dgp_ops.skeleton_matrix(edges, nj)
    d_ref = dgp_ops.skeleton_distances(mu, S0)
    t_ref = dgp_ops.temporal_distances(mu)
    out = eng.potentials(mu.cuda(), edges)
    torch.cuda.synchronize()
    print("  skel err %.3g temporal err %.3g" % ((out["skel"].cpu() - d_ref).abs().max().item(),
                                                  (out["temporal"].cpu() - t_ref).abs().max().item()), flush=True)
    # sharded with halo == unsharded
    a = eng.potentials(mu[:400].cuda(), edges, halo_next=mu[400].cuda())
    b = eng.potentials(mu[400:].cuda(), edges)
    tt = torch.cat([a["temporal"], b["temporal"]]).cpu()
    print("  halo-sharded temporal bit-exact:", bool((tt == out["temporal"].cpu()).all()), flush=True)


def diag_forward(eng_factory):
    print("== forward (layer-wise vs oracle) ==", flush=True)
    nj = 4
    W = synthetic.make_weights(nj, seed=0)
    eng = eng_factory(nj)
    eng.load_weights(W)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    for (T, H, Wd) in [(2, 235, 301), (1, 470, 640)]:
        frames, _ = synthetic.make_video(T, H, Wd, nj)
        x = torch.from_numpy(frames.astype(np.float32))
        ep = {}
        with torch.no_grad():
            net = pose_net.extract_features(x, Wt, ep)
            pred = pose_net.prediction_layer(net, Wt, "part_pred")
            loc = pose_net.prediction_layer(net, Wt, "locref_pred")
        eng.keep_activations(True)
        t0 = time.time()
        logits, locref = eng.forward(torch.from_numpy(frames).cuda())
        torch.cuda.synchronize()
        print("  forward %dx%dx%d took %.1f ms (first call incl. plan)" % (T, H, Wd, 1e3 * (time.time() - t0)))
        for name, ref in ep.items():
            try:
                got = torch.from_numpy(eng.get_activation(name))
                if got.shape != ref.shape:
                    print("  %-50s SHAPE got %s ref %s" % (name, tuple(got.shape), tuple(ref.shape)))
                    continue
                report_err(name.replace("resnet_v1_50/", ""), got, ref)
            except Exception as ex:
                print("  %-50s missing (%s)" % (name, ex))
        report_err("part_pred logits", logits, pred)
        report_err("locref", locref, loc)
        sig_err = (torch.sigmoid(logits.cpu()) - torch.sigmoid(pred)).abs().max().item()
        mu_ref, _ = dgp_ops.argmax_2d_from_cm(pred, nj, 1.0, 1.0)
        out = eng.softargmax(logits, locref)
        print("  sigmoid scoremap max abs err %.4g ; soft-argmax err (px, scoremap units) %.4g" % (
            sig_err, (out["mu"].cpu() - mu_ref).abs().max().item()), flush=True)
        eng.keep_activations(False)
    eng.close()


def main():
    sections = sys.argv[1:] or ["conv", "softargmax", "potentials", "forward"]
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    eng = E.Engine(4)
    print("SMs", eng.num_sms())
    for s in sections:
        try:
            if s.startswith("conv"):
                diag_conv(eng, int(s.split(":")[1]) if ":" in s else None)
            elif s == "softargmax":
                diag_softargmax(eng)
            elif s == "potentials":
                diag_potentials(eng)
            elif s == "forward":
                diag_forward(lambda nj: E.Engine(nj))
        except Exception:
            print("SECTION %s FAILED" % s)
            traceback.print_exc()
    print("launches", eng.launch_count())


if __name__ == "__main__":
    main()
