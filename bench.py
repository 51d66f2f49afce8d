#!/usr/bin/env python
"""bench.py -- images/sec of one COCO-AttnGAN 256x256 G+D training step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mog|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one synthetic batch (config 5 of BASELINE.json:
B=32/GPU, T=18, GF 48, DF 96, R_NUM 3): G_NET forward; D_NET64/128/256 loss + backward + Adam;
generator adversarial + KL loss, backward through the three discriminators and G, Adam, EMA
(code/coco/attngan/trainer.py:294-342), including the DAMSM words/sentence loss through the frozen Inception-v3 image
encoder (random-init stand-in weights: there is no network for the ImageNet checkpoint).  `--no-damsm` times the G+D-only
step; `config.workload` says which.

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same
through the public trainer API with pinned-host inputs copied in and the losses read back every
step; `roofline` = the dominant convolution kernel timed alone with CUDA events; `cpu_baseline`
= the oracle (CPU port of the reference) on a bounded sample.  `--impl reference` times that CPU
port alone, with all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

CFG5 = dict(GF_DIM=48, DF_DIM=96, Z_DIM=100, R_NUM=3, EMBEDDING_DIM=256, T=18)
# forward GMAC per image from SURVEY.md section 8(a.1)/8(d) (hook-counted on the reference modules),
# G+D-only step: G fwd 25.09; D step fwd 18.70 + bwd 36.86; G step D fwd 9.18 + D dgrad 9.18 + G bwd 50.18
# with the DAMSM term (Inception-v3 image encoder fwd + dgrad, words/sentence losses): 161.0 GMAC
GMAC_PER_IMAGE_GD = 149.2
GMAC_PER_IMAGE_FULL = 161.0
WORKLOAD_GD = "attngan256-coco-config5 G+D step (G fwd; 3x D loss+bwd+Adam; G adv+KL loss+bwd+Adam+EMA); no DAMSM/Inception"
WORKLOAD_FULL = ("attngan256-coco-config5 full step (G fwd; 3x D loss+bwd+Adam; G adv + DAMSM words/sentence loss through the "
                 "frozen Inception-v3 image encoder + KL, bwd, Adam, EMA)")


def workload(args):
    return WORKLOAD_GD if args.no_damsm else WORKLOAD_FULL


def gmac(args):
    return GMAC_PER_IMAGE_GD if args.no_damsm else GMAC_PER_IMAGE_FULL


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def set_cfg():
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    reset_cfg()
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM, cfg.GAN.R_NUM = CFG5["GF_DIM"], CFG5["DF_DIM"], CFG5["Z_DIM"], CFG5["R_NUM"]
    cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM = CFG5["EMBEDDING_DIM"], CFG5["T"]
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3, cfg.TRAIN.SMOOTH.LAMBDA = 4.0, 5.0, 10.0, 50.0
    return cfg


# --------------------------------------------------------------------------------------------
# Reference arm / cpu_baseline: the UNMODIFIED reference modules (baseline/_ref, staged by __graft_entry__.build()) through
# baseline/ref_harness.py -- trainer.py:294-342 with the reference's G_NET / D_NET64/128/256 / CNN_ENCODER, losses,
# optim.Adam(betas=(0.5, 0.999)) and EMA -- on the host cores.  The oracle port is the fall-back when the staged sources
# are missing (kind "port").
# --------------------------------------------------------------------------------------------
CPU_S_PER_IMAGE = 0.40      # measured on the gpurun host (16 threads): 2.7 images/s for the full step


def ref_sample_batch(steps, warmup, batch, budget_s=200.0):
    """Bounded sample: the largest power-of-two batch <= the workload's whose (steps + warmup) CPU steps fit the budget."""
    b = batch
    while b > 2 and (steps + warmup) * b * CPU_S_PER_IMAGE > budget_s:
        b //= 2
    return b


def cpu_reference_run(steps, warmup, batch, damsm=True):
    from baseline import ref_harness as H
    cores = os.cpu_count() or 1
    if H.available():
        r = H.time_attngan(batch, steps, warmup, "cpu", damsm, threads=cores)
        return {"value": r["images_per_s"], "unit": "images/s", "cores": cores, "kind": "reference",
                "sample": "unmodified reference modules (baseline/_ref: G_NET, D_NET64/128/256, CNN_ENCODER, miscc/losses.py) "
                          "driven as trainer.py:294-342 incl. optim.Adam + EMA%s, torch CPU fp32, %d threads, config 5 at a "
                          "bounded batch of %d, %d warm-up + %d timed steps"
                          % (" and the DAMSM / Inception-v3 branch" if damsm else "", cores, batch, warmup, steps),
                "ms_per_step": r["ms_per_step"], "batch": batch}
    return cpu_port_run(steps, warmup, batch, damsm)


def cpu_port_run(steps, warmup, batch, damsm=True):
    from mog_b200 import synth
    from oracle import attngan_oracle as O
    from mog_b200.attngan import model as M
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    set_cfg()
    torch.manual_seed(1234)
    sdG = synth.fill_state_dict(M.G_NET().state_dict(), 1)
    sdDs = [synth.fill_state_dict(c().state_dict(), 2 + i) for i, c in enumerate((M.D_NET64, M.D_NET128, M.D_NET256))]
    PG, PDs = O.leafify(sdG), [O.leafify(s) for s in sdDs]
    ocfg = O.Cfg(**{k: v for k, v in CFG5.items() if k != "T"})
    b = synth.attngan_batch(batch, T=CFG5["T"], nef=CFG5["EMBEDDING_DIM"], nz=CFG5["Z_DIM"], seed=1234)
    PE = None
    if damsm:
        PE = {k: v for k, v in synth.fill_encoder_state_dict(M.CNN_ENCODER(CFG5["EMBEDDING_DIM"]).state_dict(), 9).items()}
    state = O.make_train_state(PG, PDs)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(PG, PDs, state, ocfg, b, PE=PE)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return {"value": batch * len(times) / tot, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "oracle/attngan_oracle.train_step (torch CPU fp32 port of trainer.py:294-342 incl. Adam + EMA%s) at config 5, "
                      "batch %d, %d warm-up + %d timed steps (reference sources not staged on this box)"
                      % (" and DAMSM / Inception-v3" if damsm else "", batch, warmup, len(times)),
            "ms_per_step": 1e3 * tot / len(times), "batch": batch}


def workload_config(args, ws):
    """The workload both arms run (identical dict in both JSON lines); arm-specific details travel outside `config`."""
    return {"workload": workload(args), "batch_per_gpu": args.batch, "global_batch": ws * args.batch, "words": CFG5["T"],
            "l2": "working set per step (>5 GB) exceeds the 126 MB L2; no flush needed",
            "optimizer": "Adam(2e-4, betas=(0.5, 0.999)) x4 + EMA of G inside the timed region",
            "algorithmic_gflop_per_image": 2 * gmac(args)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    bs = args.ref_batch or ref_sample_batch(args.steps, args.warmup, args.batch)
    r = cpu_reference_run(args.steps, args.warmup, bs, not args.no_damsm)
    line = {"impl": "reference", "metric": "images/sec (G+D fwd+bwd) COCO-AttnGAN 256^2", "value": r["value"],
            "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": workload_config(args, ws),
            "arm": {"device": "cpu", "threads": r["cores"], "sample_batch_per_step": bs,
                    "note": "images/s of the same step on a bounded batch (CPU time per image is flat in the batch size)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def subprocess_json(cmd, timeout):
    """Run a helper (reference timing) in its own process -- the harness rebinds torch globals -- and parse its JSON line."""
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=timeout, cwd=ROOT).stdout
        for ln in reversed(out.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}
    return {"error": "no output"}


# --------------------------------------------------------------------------------------------
# libmog arm
# --------------------------------------------------------------------------------------------
def run_mog(args):
    import __graft_entry__ as ge
    from mog_b200 import _lib, ops, parallel, synth
    from mog_b200.attngan.miscc.utils import weights_init
    from mog_b200.attngan.trainer import condGANTrainer
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the libmog path has no CPU fallback "
                         "(use --impl reference for the CPU port)")
    ws = parallel.init_from_env("nccl")
    rank = parallel.rank()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        ge.build()
    if ws > 1:
        dist.barrier()
    cfg = set_cfg()
    ops.set_precision(args.precision)
    cfg.TRAIN.BATCH_SIZE = args.batch
    B, K, W = args.batch, args.steps, args.warmup

    torch.manual_seed(1234)
    tr = condGANTrainer("", None, 0, None)
    enc = None
    if not args.no_damsm:
        # frozen DAMSM image encoder in eval mode (trainer.py:56-77); deterministic stand-in for the ImageNet weights
        from mog_b200.attngan.model import CNN_ENCODER
        enc = CNN_ENCODER(CFG5["EMBEDDING_DIM"])
        enc.load_state_dict(synth.fill_encoder_state_dict(enc.state_dict(), 9))
        for p in enc.parameters():
            p.requires_grad = False
        enc.to(dev).eval()
    _, _, netG, netsD, _ = tr.build_models(image_encoder=enc, load_encoders=False)   # weights_init (orthogonal) on device, broadcast from rank 0
    optG, optDs = tr.define_optimizers(netG, netsD)
    st = tr.make_step_state(netG, netsD, optG, optDs)

    host = synth.attngan_batch(B, T=CFG5["T"], nef=CFG5["EMBEDDING_DIM"], nz=CFG5["Z_DIM"], seed=1234 + rank)
    keys = ["sent_emb", "words_embs", "mask", "transf_matrices", "transf_matrices_inv", "label_one_hot", "noise"]
    pinned = {k: host[k].pin_memory() for k in keys}
    pinned_imgs = [t.pin_memory() for t in host["imgs"]]
    h2d_bytes = sum(t.numel() * t.element_size() for t in pinned.values()) + \
        sum(t.numel() * t.element_size() for t in pinned_imgs)

    def upload():
        d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        imgs = [t.to(dev, non_blocking=True) for t in pinned_imgs]
        return d, imgs

    d, imgs = upload()

    def step(d, imgs, fresh_noise=True):
        return tr.train_step(st, imgs, d["sent_emb"], d["words_embs"], d["mask"], d["transf_matrices"],
                             d["transf_matrices_inv"], d["label_one_hot"], host["cap_lens"], host["class_ids"],
                             noise=None if fresh_noise else d["noise"])

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, n):
        """n calls bracketed by barrier+sync, CUDA events on the launching stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        host_ms[0] = 1e3 * (time.perf_counter() - t0) / n   # time the host needs to ENQUEUE one step (no sync inside)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if ws > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    gs = None
    if not args.no_graph:
        # the whole step as ONE CUDA graph (condGANTrainer.graphed_step): its eager warm-up steps count as warm-up
        gs = tr.graphed_step(st, imgs, d["sent_emb"], d["words_embs"], d["mask"], d["transf_matrices"], d["transf_matrices_inv"],
                             d["label_one_hot"], host["cap_lens"], host["class_ids"], warmup=max(1, min(W, 2)))
        for _ in range(max(0, W - 2)):
            gs.replay()
        with ClockSampler(local) as cs:
            ms = timed(gs.replay, K)
        launches = gs.launches * K
    else:
        for _ in range(W):
            step(d, imgs)
        n0 = _lib.launch_count()
        with ClockSampler(local) as cs:
            ms = timed(lambda: step(d, imgs), K)
        launches = _lib.launch_count() - n0
    clocks = cs.summary()
    host_enqueue_ms = host_ms[0]
    value = ws * B * K / (ms / 1e3)

    # ---- end to end: pinned host inputs copied in, losses read back, every step
    loss_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def e2e_step():
        if gs is not None:   # pinned host -> the graph's static input buffers, replay, losses back
            eD, eG, kl = gs(pinned_imgs, pinned["sent_emb"], pinned["words_embs"], pinned["mask"], pinned["transf_matrices"],
                            pinned["transf_matrices_inv"], pinned["label_one_hot"], host["cap_lens"], host["class_ids"])
        else:
            dd, ii = upload()
            eD, eG, kl = step(dd, ii)
        loss_host.copy_(torch.stack((eD, eG, kl)), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    ms_e2e = timed(e2e_step, K)
    e2e_value = ws * B * K / (ms_e2e / 1e3)

    line = None
    roof = cpu = None
    if rank == 0:
        roof = roofline_probe(dev, args, B)
    cudnn = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline:
        # the reference's own step on the host cores: bounded sample (batch 8, 1 warm-up + 2 timed steps), own process
        py = sys.executable
        extra = ["--no-damsm"] if args.no_damsm else []
        ref = subprocess_json([py, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                               "--ref-batch", "8", "--batch", str(B)] + extra, 600)
        cpu = ref.get("cpu_baseline") or {"error": ref.get("error", "failed")}
        # reported, not the headline: the same reference modules on this B200 under stock torch + cuDNN (cudnn.benchmark as in
        # trainer.py:51), with torch's default TF32 convolutions and with TF32 disabled
        cudnn = {}
        for tag, flag in (("tf32", "1"), ("fp32", "0")):
            r = subprocess_json([py, os.path.join(ROOT, "baseline", "ref_harness.py"), "--device", "cuda", "--batch", str(B),
                                 "--steps", "5", "--warmup", "3", "--tf32", flag] + extra, 600)
            cudnn[tag] = {"images_per_s": r.get("images_per_s"), "ms_per_step": r.get("ms_per_step")} if "error" not in r else r
    if rank == 0:
        pk, src = peaks()
        line = {"metric": "images/sec (G+D fwd+bwd) COCO-AttnGAN 256^2", "value": value, "unit": "images/s",
                "n_gpus": ws, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "fp32", "bf16x3": "bf16x3 (3-pass split, fp32-equivalent) + fp32 accumulate",
                          "bf16": "bf16 operands, fp32 accumulate"}[args.precision],
                "data": "synthetic",
                "config": workload_config(args, ws),
                "arm": {"parallelism": "dp%d (one process per GPU, NCCL grad all-reduce per net)" % ws, "precision": args.precision,
                        "optimizer_kernel": "fused libmog Adam + EMA (mog_adam_multi_dev)", "cuda_graph": gs is not None,
                        "step_tflops_achieved": 2 * gmac(args) * 1e9 * value / 1e12 / ws, "built": ge.BUILD_MODE},
                "clocks": clocks, "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 12, "ms_per_step": ms_e2e / K},
                "roofline": roof, "cpu_baseline": cpu, "torch_cudnn_b200": cudnn, "peaks": src}
        if roof is not None:
            # whole-step fraction: algorithmic FLOPs of the step / step time / SUSTAINED bf16 peak (kernels timed inside a long step)
            sus = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
            roof["step_tflops"] = 2 * gmac(args) * 1e9 * value / 1e12 / ws
            roof["step_frac"] = roof["step_tflops"] / sus
            roof["step_peak"] = sus
        print(json.dumps(line))
    if ws > 1:
        # Captured CUDA graphs hold NCCL kernels: tearing the communicator down under them hung (observed: 10 minutes, killed).
        # Everything is finished and printed -- synchronise, meet at a barrier, and leave without the NCCL teardown.
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def ncu_traffic(B):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant conv's launches, read from the committed `ncu --set full`
    capture (profiles/*conv_halo_up*.raw.csv, taken at B = 32 with the same ops.conv2d call); None when no capture is there."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_conv_halo_up*.raw.csv")))
    if not files:
        return None, None
    path = files[-1]
    try:
        rows = list(csv.reader(open(path)))
        hdr = rows[0]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        units = rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot, n = 0.0, 0
        for r in rows[2:]:
            if "conv_halo" in r[ik] and n < FWD_LAUNCHES_IN_CAPTURE:
                tot += float(r[ir].replace(",", "")) * scale.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * scale.get(units[iw], 1.0)
                n += 1
        return tot * (B / 32.0), os.path.relpath(path, ROOT) + " (first %d conv_halo launches = the forward)" % n
    except Exception as e:  # noqa: BLE001
        return None, "unreadable capture %s: %s" % (os.path.basename(path), e)


FWD_LAUNCHES_IN_CAPTURE = 4     # the forward of up2x + 3x3 is four sub-pixel phase launches in the r1c capture


def roofline_probe(dev, args, B):
    """The dominant kernel timed alone: the largest conv launch of the step, G.h_net3.upsample
    (nearest x2 + conv3x3 96->96 at 256x256; implicit GEMM M=B*65536, N=96, K=864 -- SURVEY 8(a.1)),
    forward, through the same ops.conv2d entry the model uses."""
    from mog_b200 import ops
    pk, src = peaks()
    x = torch.randn(B, 128, 128, 96, device=dev)
    w = torch.randn(96, 96, 3, 3, device=dev) * 0.03
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    with torch.no_grad():
        for _ in range(3):
            ops.conv2d(x, w, None, 1, 1, True, 0)
        times = []
        for _ in range(5):
            flush.zero_()                      # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            ops.conv2d(x, w, None, 1, 1, True, 0)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    flops = 2.0 * B * 256 * 256 * 96 * 864
    achieved = flops / (ms / 1e3) / 1e12
    peak = pk["bf16_tflops"]
    traffic, traffic_src = ncu_traffic(B)
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "executed_tflops": achieved * 3.0 / 2.25,
            "note": "achieved counts the dense fp32-equivalent FLOPs of the upsampled 3x3 conv; the kernel executes 2.25x fewer "
                    "MACs (sub-pixel phases) x 3 bf16 passes (hi*hi + lo*hi + hi*lo)",
            "kernel": "conv fwd G.h_net3.upsample (up2x + 3x3, 96->96 @256^2), precision=%s" % args.precision,
            "ms_per_launch": ms, "peak_source": src + " bf16 burst (kernel timed alone)",
            "algorithmic_flops_per_launch": flops}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mog", choices=["mog", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=0, help="bounded CPU sample: images per CPU step (0 = sized to the time budget)")
    ap.add_argument("--precision", default=os.environ.get("MOG_PRECISION", "bf16x3"), choices=["fp32", "bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph of the step")
    ap.add_argument("--no-damsm", action="store_true", help="time the G+D-only step (no DAMSM loss / Inception-v3 encoder)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "mog":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_mog(args)


if __name__ == "__main__":
    main()
