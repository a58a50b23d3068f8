"""SURVEY.md §8 f2 — the frozen DistilBERT sentence encoder (r3m/models/models_language.py:13-35) on the library's own
kernels, against transformers' DistilBertModel (the module the reference instantiates) in strict fp32 on the same GPU.
distilbert-base-uncased's weights are unreachable offline, so both sides load the same seeded random-init checkpoint of
the real architecture (full size) — the arithmetic under test does not depend on the values.  Tolerance 1e-4 relative,
ten times tighter than the north star's embedding bound: the Linears run on tf32 tensor cores with every fp32 operand
split into two tf32 terms (three products, fp32 accumulation), which measured 1.3e-3 with plain tf32 rounding."""
import pytest
import torch

from gpu_common import rel, strict_fp32

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _hf_model(seed, **cfg):
    from transformers import DistilBertConfig, DistilBertModel

    torch.manual_seed(seed)
    m = DistilBertModel(DistilBertConfig(**cfg)).eval()
    # LayerNorm weights / biases and Linear biases away from their 1 / 0 init, so that every parameter matters
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, v in m.state_dict().items():
            if k.endswith("bias"):
                v.copy_(0.05 * torch.randn(v.shape, generator=g))
            elif "LayerNorm.weight" in k or "layer_norm.weight" in k:
                v.copy_(1.0 + 0.1 * torch.randn(v.shape, generator=g))
            elif k.endswith("lin.weight") or k.endswith("lin1.weight") or k.endswith("lin2.weight"):
                v.copy_(0.05 * torch.randn(v.shape, generator=g))  # pretrained-scale weights (init std 0.02 is tame)
    return m


def _batch(seed, B, T, vocab, ragged=True):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (B, T), generator=g)
    mask = torch.ones(B, T, dtype=torch.long)
    if ragged:
        lens = torch.randint(2, T + 1, (B,), generator=g)
        lens[0] = T
        for b in range(B):
            mask[b, lens[b]:] = 0
            ids[b, lens[b]:] = 0  # [PAD]
    return ids, mask


def _reference(m, ids, mask):
    strict_fp32()
    m = m.cuda()
    with torch.no_grad():
        h = m(ids.cuda(), attention_mask=mask.cuda()).last_hidden_state
    return h.mean(1), h


@pytest.mark.parametrize("B,T", [(64, 12), (5, 33), (3, 7)])
def test_full_size_encoder_matches_transformers(B, T):
    from r3m_b200.bert import DistilBertEncoder

    m = _hf_model(0)
    ids, mask = _batch(B * 100 + T, B, T, m.config.vocab_size)
    want, want_h = _reference(m, ids, mask)
    enc = DistilBertEncoder(m.state_dict(), "cuda", max_tokens=1024)
    got, got_h = enc.encode(ids, mask, return_hidden=True)
    torch.cuda.synchronize()
    assert rel(got_h, want_h) < TOL, rel(got_h, want_h)
    assert rel(got, want) < TOL, rel(got, want)
    assert float((got_h - want_h).abs().max()) < 2e-3  # no single token is off (LayerNorm output is O(1))
    assert enc.launches_last_call == 1 + 6 * 12 + 1  # per layer: q, k, v, attention, out, LN, lin1 x 3 column slices, split, lin2, LN


def test_long_sentences_take_the_tiled_attention_path():
    """T > 64 keys: several key tiles, online softmax across them; small architecture (2 layers, 4 heads)."""
    from r3m_b200.bert import DistilBertEncoder

    cfg = dict(vocab_size=1000, max_position_embeddings=256, dim=256, n_heads=4, n_layers=2, hidden_dim=1024)
    m = _hf_model(3, **cfg)
    ids, mask = _batch(7, 3, 150, 1000)
    mask[1, 60:] = 0  # a sentence whose second and third key tiles are all padding
    want, want_h = _reference(m, ids, mask)
    enc = DistilBertEncoder(m.state_dict(), "cuda", max_tokens=512, vocab=1000, max_pos=256, dim=256, heads=4, layers=2,
                            ffn=1024)
    got, got_h = enc.encode(ids, mask, return_hidden=True)
    assert rel(got_h, want_h) < TOL, rel(got_h, want_h)
    assert rel(got, want) < TOL


def test_encoder_is_deterministic_and_mask_sensitive():
    from r3m_b200.bert import DistilBertEncoder

    m = _hf_model(1)
    ids, mask = _batch(11, 8, 16, m.config.vocab_size)
    enc = DistilBertEncoder(m.state_dict(), "cuda", max_tokens=256)
    a = enc.encode(ids, mask).clone()
    b = enc.encode(ids, mask)
    assert torch.equal(a, b)
    c = enc.encode(ids, torch.ones_like(mask))  # un-masking the padding must change ragged sentences only
    changed = (a - c).abs().amax(1) > 1e-4
    ragged = mask.sum(1) < mask.shape[1]
    assert torch.equal(changed.cpu(), ragged)


class _WhitespaceTokenizer:
    """Stands in for AutoTokenizer (vocabulary files are unreachable offline): [CLS] words [SEP], right padding."""

    def __init__(self, vocab):
        self.vocab = vocab

    def __call__(self, sentences, return_tensors="pt", padding=True):
        rows = [[101] + [1000 + (sum(map(ord, w)) % (self.vocab - 1000)) for w in s.split()] + [102] for s in sentences]
        T = max(len(r) for r in rows)
        ids = torch.tensor([r + [0] * (T - len(r)) for r in rows])
        mask = torch.tensor([[1] * len(r) + [0] * (T - len(r)) for r in rows])
        return {"input_ids": ids, "attention_mask": mask}


def test_lang_encoder_module_matches_the_reference_pipeline():
    """r3m_b200.model.LangEncoder (the reference's class, models_language.py:13-35) with an injected tokenizer /
    checkpoint: sentences -> ids -> native encoder == transformers on the same ids; state_dict keeps the reference's
    `model.*` keys and a reload refreshes the native copy."""
    from r3m_b200.model import LangEncoder

    m = _hf_model(2)
    tok = _WhitespaceTokenizer(m.config.vocab_size)
    le = LangEncoder("cuda", tokenizer=tok, hf_model=m)
    assert le.lang_size == 768
    sentences = ["open the top drawer", "pick up the red mug and put it on the shelf", "push"]
    got = le(sentences)
    enc = tok(sentences)
    want, _ = _reference(m, enc["input_ids"], enc["attention_mask"])
    assert got.shape == (3, 768) and rel(got, want) < TOL
    assert any(k.startswith("model.transformer.layer.5.") for k in le.state_dict())
    other = _hf_model(5)
    le.load_state_dict({"model." + k: v for k, v in other.state_dict().items()})
    want2, _ = _reference(other, enc["input_ids"], enc["attention_mask"])
    assert rel(le(sentences), want2) < TOL
