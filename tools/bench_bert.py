"""Sentence encoder: r3m_b200.bert.DistilBertEncoder vs transformers' DistilBertModel on the same GPU,
same random-init distilbert-base architecture, 64 sentences per call (one per clip of a c3 step).  Device-timed.
-> gpurun_out/r2_distilbert.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def timed(fn, iters=30, warm=5):
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


def main():
    from transformers import DistilBertConfig, DistilBertModel

    from r3m_b200.bert import DistilBertEncoder

    torch.manual_seed(0)
    hf = DistilBertModel(DistilBertConfig()).eval().cuda()
    enc = DistilBertEncoder(hf.state_dict(), "cuda", max_tokens=4096)
    out = {}
    for B, T in ((64, 12), (64, 24), (64, 48)):
        ids = torch.randint(0, 30522, (B, T), device="cuda")
        mask = torch.ones(B, T, dtype=torch.long, device="cuda")
        mask[:, T * 2 // 3:] = (torch.arange(B, device="cuda")[:, None] % 2 == 0).long()
        rec = {}
        rec["ours_ms"] = timed(lambda: enc.encode(ids, mask))
        rec["launches"] = enc.launches_last_call
        with torch.no_grad():
            torch.backends.cuda.matmul.allow_tf32 = False
            rec["transformers_fp32_ms"] = timed(lambda: hf(ids, attention_mask=mask).last_hidden_state.mean(1))
            want = hf(ids, attention_mask=mask).last_hidden_state.mean(1)
            torch.backends.cuda.matmul.allow_tf32 = True
            rec["transformers_tf32_ms"] = timed(lambda: hf(ids, attention_mask=mask).last_hidden_state.mean(1))
            got32 = hf(ids, attention_mask=mask).last_hidden_state.mean(1)
            torch.backends.cuda.matmul.allow_tf32 = False
        got = enc.encode(ids, mask)
        rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())  # noqa: E731
        rec["ours_rel_err_vs_fp32"] = rel(got, want)
        rec["transformers_tf32_rel_err_vs_fp32"] = rel(got32, want)
        out[f"B{B}_T{T}"] = rec
        print(B, T, rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_distilbert.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
