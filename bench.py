"""Headline benchmark: R3M pretrain-step frames/s (224x224, ResNet-50), BASELINE.json config c3 per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5                       # our arm (sm_100a engine)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1      # CPU arm: the reference algorithm on host cores

One JSON line on stdout (rank 0).  A "step" is one ``Trainer.update``: forward over 64 clips x 5 frames, L1/L2 + TCN
(+ language) losses, backward, Adam.  ``value`` is device-timed with the frames resident in HBM; ``e2e`` is the same
step through the public API starting from pinned HOST frames (H2D inside the timed region, metrics read back).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pretrain-step frames/s (224x224, ResNet-50)"
HYPER = dict(l2weight=1e-5, l1weight=1e-5, tcnweight=1.0, lr=1e-4, hidden_dim=1024)  # README.md:32 + config_rep.yaml
CLIPS_PER_GPU = 64          # BASELINE.json configs[2] (c3) / configs[4] (c5: 512 global over 8 GPUs)
TRAIN_GFLOP_PER_FRAME = 24.287  # fwd + dgrad + wgrad conv FLOPs, ResNet-50


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=50)
    ap.add_argument("--clips", type=int, default=CLIPS_PER_GPU)
    ap.add_argument("--lang", type=int, default=1, help="language head on (c3) / off")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the side measurements of BASELINE configs 0/1/3")
    ap.add_argument("--e2e-debug", default="", help="diagnostic: 'nocopy' skips the H2D copy in the e2e loop")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line): one
    streaming `nvidia-smi -lms 100` process, started before and stopped after the timed steps."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.t0 = index, None, 0.0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.35)  # first sample is out before the timed region starts
        except Exception:  # noqa: BLE001 - sampling is best effort
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:  # noqa: BLE001
                self.proc.kill()
                out = ""
            for line in out.splitlines():
                parts = [p.strip() for p in line.split(",")]
                if len(parts) == 7:
                    rows.append(parts)

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None

        sm = sorted(v for v in (num(r[0]) for r in rows) if v is not None)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        power = [v for v in (num(r[2]) for r in rows) if v is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": num(rows[0][1]) if rows else None,
                "power_w_max": max(power) if power else None,
                "samples": len(rows), "reasons": sorted(reasons)}


class StubLangEncoder:
    """Frozen-sentence-encoder stand-in (DistilBERT weights are unreachable offline): a seeded
    [B,768] embedding, identical on both arms, produced once per step like the real encoder would be."""
    lang_size = 768

    def __init__(self, device):
        self.device = device
        self._cache = {}

    def __call__(self, sentences):
        import torch

        n = len(sentences)
        if n not in self._cache:
            g = torch.Generator().manual_seed(1234)
            self._cache[n] = torch.randn(n, 768, generator=g)
        return self._cache[n]


def sentences_for(n):
    return ["" if i % 10 == 9 else "C does something %d" % i for i in range(n)]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "hbm_gbs": p.get("hbm_gbs"),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)"}
    return {"tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md"}


# ----------------------------------------------------------------------------------------------------------------
def cpu_step_runner(size, clips, lang):
    """The reference algorithm (oracle port of R3M.forward + Trainer.update) on host cores, all threads."""
    import torch
    from oracle import r3m_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    params, buffers = O.init_state(size, 0, lang=lang)
    frames = O.synthetic_frames(clips, 1)
    perms = O.draw_permutations(clips, 2)
    hyper = dict(l2weight=HYPER["l2weight"], l1weight=HYPER["l1weight"], tcnweight=HYPER["tcnweight"],
                 langweight=1.0 if lang else 0.0, lr=HYPER["lr"])
    lang_emb = O.stub_lang_embedding(clips, 3) if lang else None
    mask = torch.tensor([1.0 * (s != "") for s in sentences_for(clips)]) if lang else None
    opt = O.new_opt_state()

    def step():
        O.update(params, buffers, opt, frames, perms, hyper, size, lang_emb, mask)

    return step


CPU_SAMPLE_CLIPS = 8  # bounded sample of the 64-clip workload: 8 clips = 40 frames per step (~10-25 s of host work in all)


def cpu_baseline(size, lang, clips=CPU_SAMPLE_CLIPS, steps=2):
    step = cpu_step_runner(size, clips, lang)
    step()
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = (time.time() - t0) / steps
    import torch

    return {"value": clips * 5 / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} timed Trainer.update steps of {clips} clips ({clips * 5} frames) after 1 warm-up, "
                      f"oracle port (fp32, torch CPU ops), {os.cpu_count()} host CPUs visible"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lang = bool(args.lang)
    clips = CPU_SAMPLE_CLIPS
    step = cpu_step_runner(args.size, clips, lang)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / max(args.steps, 1)
    import torch

    v = clips * 5 / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args), reference_sample_clips_per_step=clips),
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"each step = one Trainer.update of {clips} clips ({clips * 5} frames): the "
                                       "reference algorithm (oracle port; /root/reference is pure Python whose "
                                       "arithmetic is torch CPU ops) on all host threads"},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def family_traffic(args, clips, family):
    """DRAM bytes per launch of a kernel family (dram__bytes_read.sum + dram__bytes_write.sum averaged over the family's
    launches of one step) from the committed ncu capture of this exact workload (profiles/r2_traffic.json, produced by
    tools/ncu_per_kernel.py from an `ncu` pass over tools/profile_step.py); None when no capture of this workload /
    family is committed — DRAM counters cannot be read outside a profiler, and a number taken under one is never a
    bench value, so the live part of the roofline is `achieved` (CUDA events) and this is the profiler's part."""
    if args.size != 50 or clips != 64 or not args.lang:
        return None, None
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f)["families"][family]
        return rec["dram_bytes_per_launch"], "profiles/r2_traffic.json (ncu, same workload)"
    except (OSError, KeyError, ValueError):
        return None, None


FAMILY_KERNELS = {
    "conv_igemm": "conv_igemm_kernel (tcgen05 implicit GEMM: forward convs + dgrad)",
    "wgrad": "wgrad_kernel (tcgen05 filter gradients, fp32 split-K)",
    "norm": "BatchNorm family: bn_apply / bn_bwd_reduce / bn_bwd_apply / stem_bwd / preprocess_stem (HBM streaming)",
    "pool": "stem BN+ReLU+maxpool, average pool",
    "loss": "fused LP / TCN loss heads",
    "optim": "fused Adam + filter re-packs",
    "lang": "language-reward head (fp32 SIMT GEMM chain)",
}
TENSOR_FAMILIES = ("conv_igemm", "wgrad", "lang")


def family_entry(name, f, pk, total_ms):
    tensor = name in TENSOR_FAMILIES
    ms = f["ms"]
    achieved = None
    if ms > 0:
        achieved = (f["flops"] / (ms * 1e-3) / 1e12) if tensor else (f["bytes"] / (ms * 1e-3) / 1e9)
    peak = pk["tflops"] if tensor else pk["hbm_gbs"]
    return {"kernel": FAMILY_KERNELS.get(name, name), "family": name, "bound": "tensor" if tensor else "hbm",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s" if tensor else "GB/s",
            "frac": (achieved / peak) if achieved else None, "launches_per_step": f["launches"],
            "avg_launch_ms": ms / max(f["launches"], 1), "share_of_step": ms / total_ms if total_ms else None,
            "algorithmic_gbytes_per_step": f["bytes"] / 1e9, "algorithmic_tflop_per_step": f["flops"] / 1e12}


def gpu_reference(args, clips, lang, dev):
    """The survey's bar (SURVEY.md §8d "Reference timed beside it (2)"): the reference pipeline on THIS GPU through
    torch / torchvision / cuDNN at the same config — as written (fp32 API, cudnn.benchmark, TF32 convs allowed) and in
    its strongest fair form (bf16 autocast + channels_last).  Device-timed like `value`; frames resident."""
    import torch
    from oracle import torch_reference as T

    out = {"what": "oracle/torch_reference.py: torchvision ResNet + the reference's update step, same clips per step",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "variants": {}}
    torch.cuda.empty_cache()
    for variant in ("as_written", "bf16_channels_last"):
        try:
            r = T.time_update(args.size, clips, variant, dev, lang=lang, steps=6, warmup=3)
            out["variants"][variant] = {"value": r["frames_per_s"], "unit": "frames/s", "ms_per_step": r["ms_per_step"]}
        except Exception as e:  # noqa: BLE001 - a reported side measurement must not sink the bench line
            out["variants"][variant] = {"error": repr(e)[:200]}
    T.configure("as_written")
    torch.backends.cudnn.benchmark = False
    return out


def other_configs(dev):
    """The other BASELINE.json configs, measured in the same run at N = 1 (device-timed, inputs resident; parity at these
    sizes is tests/test_fullsize_gpu.py's job): configs[0] ResNet-18 forward batch 4 (the load_r3m / example.py path, on
    the GPU), configs[1] ResNet-50 forward batch 256 (eval: BatchNorm folded; train: batch statistics), configs[3]
    ResNet-34 update() with 128 clips."""
    import torch
    import r3m_b200
    from r3m_b200 import R3M, Trainer

    def timed(fn, iters, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {}
    try:
        for key, size, batch in (("c1_resnet18_forward_b4", 18, 4), ("c2_resnet50_forward_b256", 50, 256)):
            m = R3M("cuda", HYPER["lr"], HYPER["hidden_dim"], size=size, langweight=0.0).to(dev)
            x = torch.randint(0, 255, (batch, 3, 224, 224), device=dev).float()
            rec = {}
            with torch.no_grad():
                for mode in ("eval", "train"):
                    m.eval() if mode == "eval" else m.train()
                    ms = timed(lambda: m(x), 20, 5)
                    rec[mode] = {"value": batch / ms * 1e3, "unit": "frames/s", "ms": ms}
            out[key] = rec
            del m, x
            torch.cuda.empty_cache()
        clips = 128
        r3m_b200.set_lang_encoder_factory(StubLangEncoder)
        m = R3M("cuda", HYPER["lr"], HYPER["hidden_dim"], size=34, l2weight=HYPER["l2weight"], l1weight=HYPER["l1weight"],
                langweight=1.0, tcnweight=HYPER["tcnweight"])
        model = torch.nn.DataParallel(m.to(dev), device_ids=[dev.index])
        tr = Trainer(eval_freq=10 ** 9)
        frames = torch.randint(0, 255, (clips, 5, 3, 224, 224), device=dev).float()
        sents = sentences_for(clips)
        ms = timed(lambda: tr.update(model, (frames, sents), 0), 8, 4)
        out["c4_resnet34_update_b128"] = {"value": clips * 5 / ms * 1e3, "unit": "frames/s", "ms_per_step": ms}
        del m, model, tr, frames
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001 - a reported side measurement must not sink the bench line
        out["error"] = repr(e)[:200]
    return out


def workload_config(args, clips_override=None):
    clips = clips_override or args.clips
    return {"workload": f"c3/c5: full Trainer.update(): ResNet-{args.size}, 5 frames/clip, {clips} clips per GPU, "
                        f"TCN + {'language + ' if args.lang else ''}L1/L2 losses, backward, Adam",
            "clips_per_gpu": clips, "frames_per_step_per_gpu": clips * 5, "language_head": bool(args.lang),
            "language_encoder": "stub embedding (DistilBERT weights unreachable offline; excluded on both arms)",
            "l2_policy": "per-step inputs + saved activations (>10 GB) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"dp{args.gpus}"}


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import r3m_b200
    from r3m_b200 import R3M, Trainer

    lang = bool(args.lang)
    r3m_b200.set_lang_encoder_factory(StubLangEncoder)
    torch.manual_seed(0)  # identical initial weights on every rank
    model = R3M("cuda", HYPER["lr"], HYPER["hidden_dim"], size=args.size, l2weight=HYPER["l2weight"],
                l1weight=HYPER["l1weight"], langweight=1.0 if lang else 0.0, tcnweight=HYPER["tcnweight"])
    model = torch.nn.DataParallel(model.to(dev), device_ids=[local])
    trainer = Trainer(eval_freq=10 ** 9)
    B = args.clips
    g = torch.Generator(device=dev).manual_seed(1 + rank)  # config_rep.yaml:16 seed 1 (+ rank)
    frames = torch.randint(0, 255, (B, 5, 3, 224, 224), generator=g, device=dev).float()
    b_lang = sentences_for(B)
    torch.manual_seed(100 + rank)  # permutation stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    metrics_box = {}

    def step_resident():
        metrics_box["m"], _ = trainer.update(model, (frames, b_lang), 0)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_resident, args.steps)
    clocks = sampler.stop()
    launches = trainer.last_launches
    frames_per_step = B * 5 * world
    value = frames_per_step * args.steps / (ms / 1e3)

    # ---- end to end through the public API: pinned HOST frames -> r3m_b200.FrameFeeder (H2D on a copy stream, two
    # slots: the upload of batch i+1 overlaps the step on batch i, what a pin_memory DataLoader + non_blocking .cuda()
    # gives the reference's loop, train_representation.py:102-104) -> Trainer.update -> metrics D2H, every step.
    # Headline: uint8 frames, the input pipeline's native format (r3m_b200/data.py; the reference's decoder also
    # yields uint8, data_loaders.py:30-32).  The fp32 batches the reference's loader would hand over are timed beside.
    from r3m_b200 import FrameFeeder

    def endless(host):
        while True:
            yield host, b_lang

    def time_e2e(host):
        feeder = FrameFeeder(endless(host), dev, as_uint8=False)

        def step():
            batch, lang_b = next(feeder)
            metrics_box["m"], _ = trainer.update(model, (batch, lang_b), 0)

        for _ in range(2):
            step()
        return timed(step, args.steps)

    host8 = torch.empty(frames.shape, dtype=torch.uint8).pin_memory()
    host8.copy_(frames.to(torch.uint8))
    ms_e2e = time_e2e(host8)
    small = 15 * B * 4 + (B * 769 * 4 if lang else 0)  # permutations (+ sentence embedding and mask): side-band pull
    e2e = {"value": frames_per_step * args.steps / (ms_e2e / 1e3), "unit": "frames/s",
           "h2d_bytes_per_step": int(host8.numel() + small), "d2h_bytes_per_step": 64,
           "ms_per_step": ms_e2e / args.steps,
           "api": "for frames, sentences in r3m_b200.FrameFeeder(loader, device): r3m_b200.Trainer.update("
                  "DataParallel(R3M), (frames, sentences), step) - uint8 frames from pinned host memory"}
    host32 = torch.empty(frames.shape, dtype=torch.float32).pin_memory()
    host32.copy_(frames)
    ms_e2e32 = time_e2e(host32)
    e2e["fp32_frames"] = {"value": frames_per_step * args.steps / (ms_e2e32 / 1e3), "unit": "frames/s",
                          "h2d_bytes_per_step": int(host32.numel() * 4 + small), "ms_per_step": ms_e2e32 / args.steps}
    del host32

    # ---- roofline of the dominant kernel family, measured live with in-stream CUDA events
    m = model.module
    eng = m._engine(B * 5)
    from r3m_b200.trainer import draw_permutations

    perms = draw_permutations(B, m.langweight, m.tcnweight).to(dev)
    emb = mask = None
    if lang:
        emb = m.lang_enc(b_lang).to(dev).float().contiguous()
        mask = torch.tensor([1.0 * (s != "") for s in b_lang], device=dev)
    fam = None
    for _ in range(3):
        m.encoder_opt.steps += 1
        fam = eng.profile_update(frames.reshape(-1, 3, 224, 224), perms, emb, mask, HYPER["l2weight"],
                                 HYPER["l1weight"], float(m.langweight), HYPER["tcnweight"], HYPER["lr"],
                                 m.encoder_opt.steps)
    pk = peaks()
    total_ms = sum(f["ms"] for f in fam.values())
    # the roofline subject is the family that takes the most device time in the step (not a hard-coded kernel)
    dom = max(fam, key=lambda k: fam[k]["ms"])
    roofline = family_entry(dom, fam[dom], pk, total_ms)
    roofline["peak_source"] = pk["source"]
    roofline["traffic"], roofline["traffic_source"] = family_traffic(args, B, dom)
    roofline["dominant_family"] = dom
    roofline["secondary"] = [family_entry(k, fam[k], pk, total_ms) for k in ("conv_igemm", "wgrad", "norm")
                             if k != dom and fam[k]["launches"]]
    for sec in roofline["secondary"]:
        sec["traffic"], sec["traffic_source"] = family_traffic(args, B, sec["family"])
    roofline["families"] = {k: {"ms": round(v["ms"], 4), "launches": v["launches"],
                                "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 and v["flops"] else None,
                                "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 and v["bytes"] else None}
                            for k, v in fam.items() if v["launches"]}
    # layer-wise roofline of the step: every launch at max(tensor time, HBM time) of its algorithmic work
    ops = eng.profile_ops()
    layerwise_ms = sum(max(fl / (pk["tflops"] * 1e12), by / (pk["hbm_gbs"] * 1e9)) * 1e3 for _, _, fl, by, _ in ops)
    roofline["layerwise_roofline_ms"] = layerwise_ms
    roofline["step_frac_of_layerwise"] = layerwise_ms / (ms / args.steps) if ms > 0 else None
    roofline["accounted_gbytes_per_step"] = sum(by for _, _, _, by, _ in ops) / 1e9
    roofline["whole_step_fraction_of_tensor_peak"] = (value / world * TRAIN_GFLOP_PER_FRAME / 1e3 / pk["tflops"]
                                                      if args.size == 50 else None)

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches * args.steps, "launches_per_step": launches,
            "roofline": roofline, "last_metrics": metrics_box.get("m")}
    if rank == 0:
        if world == 1 and not args.no_other_configs and args.size == 50 and B == CLIPS_PER_GPU:
            line["other_configs"] = other_configs(dev)
        if world == 1 and not args.no_gpu_reference:
            line["gpu_reference"] = gpu_reference(args, B, lang, dev)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.size, lang)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 when
    # the first communicator is created), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the
    # saved original descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
