"""``R3M`` — drop-in for the reference module (r3m/models/models_r3m.py:21-107) whose conv stack runs on the sm_100a
engine instead of torchvision/cuDNN.

What is kept byte-compatible with the reference:
  * constructor signature and the attributes ``Trainer`` reads (``l2weight, l1weight, langweight, tcnweight,
    num_negatives, encoder_opt``; models_r3m.py:22-34,76);
  * ``forward(obs, num_ims=1, obs_shape=[3,224,224]) -> float32 [N, outdim]`` (models_r3m.py:84-100);
  * ``sim`` (:102-107) and ``get_reward`` (:78-81);
  * the ``state_dict()`` key set, shapes (OIHW fp32) and dtypes of torchvision's ResNet under ``convnet.`` and of
    ``lang_rew.pred.{0,2,4,6,8}`` — reference checkpoints load, and ours load into the reference.

How it differs inside: all parameters are *views* into one flat device block (fp32 master weights in the layout the
kernels want, plus gradients and Adam moments); ``encoder_opt`` is the fused Adam of the engine.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib as L
from .engine import (KIND_CONV, KIND_LINEAR_B, KIND_LINEAR_W, KIND_RUN_MEAN, KIND_RUN_VAR, KIND_STEM, KIND_VECTOR,
                     Engine, Layout, _aligned_empty)

OUTDIM = {18: 512, 34: 512, 50: 2048}  # models_r3m.py:45,48,51
LANG_DIM = 768                          # models_language.py:21


class _Node(nn.Module):
    """Structural node of the state_dict tree (convnet.layer1.0.bn1 ...). Holds views, computes nothing."""


class _FusedAdam:
    """``encoder_opt``: torch.optim.Adam(params, lr) (models_r3m.py:76) executed by the engine's fused kernel over the
    flat parameter block.  Exposes the calls ``Trainer.update`` makes (trainer.py:156-158)."""

    def __init__(self, model, lr):
        self._model = model
        self.param_groups = [{"lr": lr, "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0}]
        self.steps = 0

    def zero_grad(self, set_to_none=False):
        self._model._flat(1).zero_()

    def step(self, grad_scale=1.0):
        eng = self._model._any_engine()
        self.steps += 1
        eng.adam_step(float(self.param_groups[0]["lr"]), float(grad_scale), self.steps)

    def state_dict(self):
        return {"step": self.steps, "exp_avg": self._model._flat(2).clone(), "exp_avg_sq": self._model._flat(3).clone(),
                "param_groups": self.param_groups}

    def load_state_dict(self, sd):
        self.steps = int(sd["step"])
        self._model._flat(2).copy_(sd["exp_avg"])
        self._model._flat(3).copy_(sd["exp_avg_sq"])
        self.param_groups = sd["param_groups"]


class _EncodeFn(torch.autograd.Function):
    """Train-mode ``R3M.forward`` as an autograd node, so that a host that builds the loss itself — the reference's own
    ``Trainer.update`` (r3m/trainer.py:41,155-158: ``alles = model(b_im_r)`` ... ``full_loss.backward()``) or a
    fine-tuning user of ``load_r3m`` — back-propagates through the sm_100a engine.  ``anchor`` is a parameter passed
    only so that the output requires grad; parameter gradients are written by the engine straight into the flat
    gradient region that every ``param.grad`` aliases, with autograd's accumulate semantics."""

    @staticmethod
    def forward(ctx, model, eng, x, anchor):
        out = eng.forward(x, True)
        eng._fwd_ticket = getattr(eng, "_fwd_ticket", 0) + 1
        ctx.model, ctx.eng, ctx.ticket = model, eng, eng._fwd_ticket
        return out

    @staticmethod
    def backward(ctx, dE):
        model, eng = ctx.model, ctx.eng
        if eng._fwd_ticket != ctx.ticket:
            raise L.R3MB200Error("backward through an R3M.forward whose saved activations were overwritten by a later "
                                 "forward of the same frame count (one outstanding train-mode forward per engine)")
        G = model._flat(1)
        stash = G.clone()  # whatever autograd / earlier backward passes have accumulated so far (language head included)
        G.zero_()
        eng.backward(dE.contiguous().float())
        G.add_(stash)
        return None, None, None, None


class R3M(nn.Module):
    def __init__(self, device, lr, hidden_dim, size=34, l2weight=1.0, l1weight=1.0, langweight=1.0, tcnweight=0.0,
                 l2dist=True, bs=16):
        super().__init__()
        if size not in OUTDIM:
            raise NameError("Invalid ResNet size %r: r3m_b200 builds ResNet-18/34/50 (models_r3m.py:44-52)" % (size,))
        self.device = device
        self.use_tb = False
        self.l2weight = l2weight
        self.l1weight = l1weight
        self.tcnweight = tcnweight
        self.l2dist = l2dist
        self.langweight = langweight
        self.size = size
        self.num_negatives = 3
        self.outdim = OUTDIM[size]
        self.hidden_dim = hidden_dim
        self.cs = torch.nn.CosineSimilarity(1)

        self._has_lang = langweight > 0.0
        self._layout = Layout(size, self._has_lang, hidden_dim)
        self._block = _aligned_empty(self._layout.param_block_bytes, torch.device("cpu"), zero=True)
        self._bn_names = []
        self._engines = OrderedDict()
        self.eval_precision = "bf16"  # see set_eval_precision
        self._synced_version = None
        self._dirty = True
        self._build_tree()
        self._init_weights()
        if self._has_lang:
            self.lang_enc = _make_lang_encoder(device)
        self.encoder_opt = _FusedAdam(self, lr)
        self._register_state_dict_hook(_clone_state_dict_entries)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.mark_weights_dirty())
        self.train()

    # ------------------------------------------------------------------------------------------------ storage
    def _flat(self, which):
        """fp32 view of a parameter-block region: 0 params, 1 grads, 2 Adam m, 3 Adam v, 4 BN buffers."""
        off = self._layout.region_offsets[which]
        n = self._layout.num_buffer_floats if which == 4 else self._layout.num_params
        return self._block[off:off + 4 * n].view(torch.float32)

    @staticmethod
    def _view(flat, info):
        t = flat[info.offset:info.offset + info.numel]
        if info.kind == KIND_CONV:
            co, ci, r, s = info.shape
            return t.view(co, r, s, ci).permute(0, 3, 1, 2)  # stored KRSC, presented OIHW
        return t.view(info.shape)

    def _node_for(self, dotted):
        node = self
        for part in dotted:
            if not hasattr(node, part):
                node.add_module(part, _Node())
            node = getattr(node, part)
        return node

    def _build_tree(self):
        P, G, buf = self._flat(0), self._flat(1), self._flat(4)
        nbn = sum(1 for t in self._layout.tensors if t.kind == KIND_RUN_MEAN)
        first = not hasattr(self, "_nbt")
        if first:
            self._nbt = torch.zeros(nbn, dtype=torch.long)
        ibn = 0
        for info in self._layout.tensors:
            *path, leaf = info.name.split(".")
            node = self._node_for(path)
            if info.kind in (KIND_RUN_MEAN, KIND_RUN_VAR):
                node._buffers[leaf] = self._view(buf, info)
                if info.kind == KIND_RUN_VAR:
                    node._buffers["num_batches_tracked"] = self._nbt[ibn]
                    ibn += 1
            else:
                if first:
                    node._parameters[leaf] = nn.Parameter(self._view(P, info), requires_grad=True)
                else:
                    node._parameters[leaf].data = self._view(P, info)
                node._parameters[leaf].grad = self._view(G, info)

    def _init_weights(self):
        """Reference init laws: tv resnet.py:208-213 (kaiming_normal_ fan_out/relu for convs, BN weight 1 / bias 0) and
        nn.Linear's default (kaiming_uniform_(a=sqrt(5)) == U(+-1/sqrt(fan_in)) for weight and bias)."""
        with torch.no_grad():
            for info in self._layout.tensors:
                *path, leaf = info.name.split(".")
                node = self._node_for(path)
                if info.kind in (KIND_CONV, KIND_STEM):
                    co, _ci, r, s = info.shape
                    node._parameters[leaf].normal_(0.0, math.sqrt(2.0 / (co * r * s)))
                elif info.kind == KIND_VECTOR:
                    node._parameters[leaf].fill_(1.0 if leaf == "weight" else 0.0)
                elif info.kind == KIND_RUN_VAR:
                    node._buffers[leaf].fill_(1.0)
                elif info.kind == KIND_LINEAR_W:
                    bound = 1.0 / math.sqrt(info.shape[1])
                    node._parameters[leaf].uniform_(-bound, bound)
                    bias = node._parameters["bias"]
                    bias.uniform_(-bound, bound)
                elif info.kind == KIND_LINEAR_B:
                    pass  # initialised together with its weight (same fan_in bound)

    def _apply(self, fn, recurse=True):
        """.cuda() / .to(device): move the flat block as ONE tensor and re-point every parameter view at it (the
        default per-parameter conversion would break the aliasing the engine relies on)."""
        probe = fn(torch.empty(0, dtype=torch.uint8, device=self._block.device))
        if probe.device != self._block.device:
            moved = _aligned_empty(self._block.numel(), probe.device, zero=False)
            moved.copy_(self._block)
            self._block = moved
            self._nbt = self._nbt.to(probe.device)
            self._engines.clear()
            self._build_tree()
            self.mark_weights_dirty()
        if self._has_lang and isinstance(getattr(self, "lang_enc", None), nn.Module):
            self.lang_enc._apply(fn)
        return self

    def mark_weights_dirty(self):
        """Call after modifying parameters outside load_state_dict / encoder_opt (the bf16 operands are refreshed)."""
        self._dirty = True

    # ------------------------------------------------------------------------------------------------ engines
    def _engine(self, frames):
        if not self._block.is_cuda:
            raise L.R3MB200Error("r3m_b200.R3M runs on an sm_100 GPU only: call .cuda() first (no CPU fallback)")
        eng = self._engines.get(frames)
        if eng is None:
            while len(self._engines) >= 3:
                self._engines.popitem(last=False)
            eng = Engine(self.size, frames, self._block, self._has_lang, self.hidden_dim, l2dist=self.l2dist)
            eng.set_precision(self.eval_precision)
            self._engines[frames] = eng
        else:
            self._engines.move_to_end(frames)
        version = self._flat(0)._version
        if self._dirty or version != self._synced_version:
            eng.sync_weights()
            self._dirty = False
            self._synced_version = version
        return eng

    def set_eval_precision(self, tier):
        """Precision tier of the eval-mode forward (``model.eval(); model(frames)``, the ``load_r3m`` user's path):

        * ``"bf16"`` (default): bf16 storage, fp32 accumulation — embeddings within 5e-3 of the fp32 reference;
        * ``"tf32"``: fp32 storage rounded to tf32, ``kind::tf32`` tensor cores — within 1e-3 (the north-star parity
          tolerance; the reference computes in fp32, models_r3m.py:97-99), at roughly half the throughput.

        Train-mode forwards and ``Trainer.update`` always run the bf16 tier."""
        if tier not in ("bf16", "tf32"):
            raise ValueError(tier)
        self.eval_precision = tier
        for eng in self._engines.values():
            eng.set_precision(tier)
        return self

    def _any_engine(self):
        if not self._engines:
            raise L.R3MB200Error("encoder_opt.step() before any forward/update")
        return next(reversed(self._engines.values()))

    # ------------------------------------------------------------------------------------------------ reference API
    def forward(self, obs, num_ims=1, obs_shape=[3, 224, 224]):  # noqa: B006 - reference signature
        """models_r3m.py:84-100.  obs in [0, 255], any dtype, [N, 3, H, W]."""
        x = obs
        if list(obs_shape) != [3, 224, 224]:
            x = _resize256_center_crop224(x.float())  # transforms.Resize(256) + CenterCrop(224), models_r3m.py:85-90
        elif x.dtype != torch.uint8:  # uint8 frames are consumed as they are (the engine converts in registers)
            x = x.float()
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
            raise ValueError(f"expected [N, 3, 224, 224] frames, got {tuple(x.shape)}")
        if x.device != self._block.device:  # one GPU per process: inputs always follow the parameter block
            x = x.to(self._block.device)
        x = x.contiguous()
        eng = self._engine(x.shape[0])
        if self.training and torch.is_grad_enabled():
            out = _EncodeFn.apply(self, eng, x, self._anchor())
        else:
            out = eng.forward(x, self.training)
        if self.training:
            self._nbt += 1
        return out

    def _anchor(self):
        return self.convnet.conv1._parameters["weight"]

    def _replicate_for_data_parallel(self):
        raise L.R3MB200Error("r3m_b200.R3M cannot be replicated by nn.DataParallel: its parameters are views into one "
                             "flat device block.  Run one process per GPU (torchrun) and wrap with "
                             "nn.DataParallel(model, device_ids=[torch.cuda.current_device()]) — a direct call")

    def sim(self, tensor1, tensor2):
        """models_r3m.py:102-107."""
        if self.l2dist:
            return -torch.linalg.norm(tensor1 - tensor2, dim=-1)
        return self.cs(tensor1, tensor2)

    def get_reward(self, e0, es, sentences):
        """models_r3m.py:78-81 — convenience path in torch ops on the same master weights (the fused engine path is
        what Trainer.update uses)."""
        if not self._has_lang:
            raise AttributeError("R3M was built with langweight == 0: no language head (models_r3m.py:67-72)")
        le = self.lang_enc(sentences)
        h = torch.cat([e0, es, le.to(e0.device)], -1)
        for i in range(5):
            lin = getattr(self.lang_rew.pred, str(2 * i))
            h = torch.nn.functional.linear(h, lin.weight, lin.bias)
            if i < 4:
                h = torch.relu(h)
        return h.squeeze(), {}


def _clone_state_dict_entries(module, state_dict, prefix, local_metadata):
    """state_dict() must not expose views of the flat block: torch.save would serialise the whole block."""
    for k in list(state_dict.keys()):
        if k.startswith(prefix):
            state_dict[k] = state_dict[k].detach().clone().contiguous()
    return state_dict


def _resize256_center_crop224(x):
    h, w = x.shape[-2:]
    if h <= w:
        nh, nw = 256, max(1, int(256 * w / h))
    else:
        nh, nw = max(1, int(256 * h / w)), 256
    x = torch.nn.functional.interpolate(x, size=(nh, nw), mode="bilinear", antialias=True, align_corners=False)
    top, left = int(round((nh - 224) / 2.0)), int(round((nw - 224) / 2.0))
    return x[..., top:top + 224, left:left + 224]


class LangEncoder(nn.Module):
    """models_language.py:13-35: frozen distilbert-base-uncased, mean-pooled last hidden state (padding included).
    The tokenizer is the reference's (transformers AutoTokenizer, host side); the encoder arithmetic runs on this
    library's kernels (r3m_b200.bert.DistilBertEncoder: tcgen05 tf32 Linears, fp32 attention / LayerNorm), loaded from
    the same checkpoint.  `tokenizer` / `hf_model` can be injected (offline use, tests); by default both come from
    `from_pretrained("distilbert-base-uncased")` like the reference's."""

    def __init__(self, device, finetune=False, scratch=False, tokenizer=None, hf_model=None):
        super().__init__()
        self.device = device
        self.modelname = "distilbert-base-uncased"
        if tokenizer is None or hf_model is None:
            from transformers import AutoModel, AutoTokenizer

            tokenizer = tokenizer or AutoTokenizer.from_pretrained(self.modelname)
            hf_model = hf_model or AutoModel.from_pretrained(self.modelname)
        self.tokenizer = tokenizer
        cfg = hf_model.config
        self._dims = dict(vocab=cfg.vocab_size, max_pos=cfg.max_position_embeddings, dim=cfg.dim, heads=cfg.n_heads,
                          layers=cfg.n_layers, ffn=cfg.hidden_dim)
        if getattr(cfg, "activation", "gelu") != "gelu" or getattr(cfg, "sinusoidal_pos_embds", False):
            raise ValueError("r3m_b200.LangEncoder implements DistilBERT with GELU and learned position embeddings")
        # the checkpointed module (state_dict keys `lang_enc.model.*`, like the reference's); its weights are copied into
        # the native encoder at first use and again after every load_state_dict
        self.model = hf_model
        self.lang_size = cfg.dim
        self._enc = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: setattr(module, "_enc", None))

    def _encoder(self):
        if self._enc is None:
            from .bert import DistilBertEncoder

            dev = torch.device(self.device)
            if dev.type == "cuda" and dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            self._enc = DistilBertEncoder(self.model.state_dict(), dev, **self._dims)
        return self._enc

    def forward(self, langs):
        try:
            langs = langs.tolist()
        except AttributeError:
            pass
        with torch.no_grad():
            enc = self.tokenizer(list(langs), return_tensors="pt", padding=True)
            return self._encoder().encode(enc["input_ids"], enc["attention_mask"])


_LANG_ENCODER_FACTORY = None


def set_lang_encoder_factory(factory):
    """Inject the sentence encoder (callable(device) -> module with .lang_size and __call__(list[str]) -> [B,768]).
    Needed offline, where the DistilBERT weights cannot be downloaded."""
    global _LANG_ENCODER_FACTORY
    _LANG_ENCODER_FACTORY = factory


def _make_lang_encoder(device):
    if _LANG_ENCODER_FACTORY is not None:
        return _LANG_ENCODER_FACTORY(device)
    return LangEncoder(device, 0, 0)
