"""Prefix conditioning (torch; runs once per utterance - SURVEY.md 2.1 row 12, 8(f) rank 4).

Keeps the reference's contract (`zonos/conditioning.py`): conditioner classes and parameter names
(`prefix_conditioner.conditioners.{i}.*`, `.project`, `.norm`) so checkpoints load unchanged,
`PrefixConditioner.forward(cond_dict) -> [B, Lc, D]`, `make_cond_dict(...)`, and the optional LRU cache of
`zonos/utilities/conditioning_cache.py`.  The text front end (eSpeak phonemes, number spelling) is imported lazily: it
needs `phonemizer` / `inflect`, which are host-side dependencies outside the hot path.
"""
import hashlib
import re
from collections import OrderedDict
from typing import Any, Iterable

import torch
import torch.nn as nn

from .config import PrefixConditionerConfig

# ---- phoneme vocabulary (zonos/conditioning.py:227-237; must match the checkpoint's embedding rows) ----------------
PAD_ID, UNK_ID, BOS_ID, EOS_ID = 0, 1, 2, 3
_N_SPECIAL = 4
_PUNCT = ';:,.!?¡¿—…"«»“”() *~-/\\&'
_LATIN = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
_IPA = ("ɑɐɒæɓʙβɔɕçɗɖðʤəɘɚɛɜɝɞɟʄɡɠɢʛɦɧħɥʜɨɪʝɭɬɫɮʟɱɯɰŋɳɲɴøɵɸθœɶʘɹɺɾɻʀʁɽʂʃʈʧʉʊʋⱱʌɣɤʍχʎʏʑʐʒʔʡʕʢǀǁǂǃˈˌːˑʼʴʰʱʲʷˠˤ˞↓↑→↗↘'̩'ᵻ")
SYMBOLS = [*_PUNCT, *_LATIN, *_IPA]
_SYMBOL_ID = {s: i for i, s in enumerate(SYMBOLS, start=_N_SPECIAL)}


def tokenize_phonemes(phonemes: list[str]) -> tuple[torch.Tensor, list[int]]:
    """BOS + symbol ids + EOS per utterance, LEFT-padded with PAD to the longest (zonos/conditioning.py:248-253)."""
    ids = [[BOS_ID, *(_SYMBOL_ID.get(ch, UNK_ID) for ch in p), EOS_ID] for p in phonemes]
    lengths = [len(i) for i in ids]
    longest = max(lengths)
    return torch.tensor([[PAD_ID] * (longest - len(i)) + i for i in ids]), lengths


def _spell_numbers(text: str) -> str:
    """Digits -> words where `inflect` is available (the reference's number normalisation, conditioning.py:199-221)."""
    if not re.search(r"[0-9]", text):
        return text
    try:
        import inflect
    except ImportError as e:
        # never diverge silently: without the number normalisation eSpeak would read the digits its own way and the
        # conditioning would differ from the reference's for the same text
        raise RuntimeError("text with digits needs the `inflect` package for the reference's number normalisation "
                           "(zonos/conditioning.py:199-221); spell the numbers out or install it") from e
    eng = inflect.engine()
    text = re.sub(r"([0-9][0-9,]+[0-9])", lambda m: m.group(1).replace(",", ""), text)
    text = re.sub(r"£([0-9,]*[0-9]+)", r"\1 pounds", text)
    text = re.sub(r"\$([0-9.,]*[0-9]+)", lambda m: m.group(1) + " dollars", text)
    text = re.sub(r"([0-9]+\.[0-9]+)", lambda m: m.group(1).replace(".", " point "), text)
    text = re.sub(r"[0-9]+(st|nd|rd|th)", lambda m: eng.number_to_words(m.group(0)), text)
    return re.sub(r"[0-9]+", lambda m: eng.number_to_words(m.group(0), andword=""), text)


_BACKENDS: dict = {}


def phonemize(texts: list[str], languages: list[str]) -> list[str]:
    """eSpeak-NG phonemes with stress marks and preserved punctuation (zonos/conditioning.py:291-335)."""
    try:
        from phonemizer.backend import EspeakBackend
    except ImportError as e:  # pragma: no cover - host dependency
        raise RuntimeError("text conditioning needs the `phonemizer` package and eSpeak-NG (not part of the GPU hot path); "
                           "pass precomputed prefix_conditioning to generate() instead") from e
    out = []
    for text, lang in zip(texts, languages):
        if lang not in _BACKENDS:
            _BACKENDS[lang] = EspeakBackend(lang, preserve_punctuation=True, with_stress=True, punctuation_marks=_PUNCT)
        out.append(_BACKENDS[lang].phonemize([_spell_numbers(text)], strip=True)[0])
    return out


# ---- conditioners --------------------------------------------------------------------------------------------------
class Conditioner(nn.Module):
    """zonos/conditioning.py:14-109: apply_cond -> project; `None` input -> the learned unconditional vector."""

    def __init__(self, output_dim: int, name: str, cond_dim: int | None = None, projection: str = "none",
                 uncond_type: str = "none", **kwargs):
        super().__init__()
        self.name, self.output_dim = name, output_dim
        self.cond_dim = cond_dim = cond_dim or output_dim
        if projection == "linear":
            self.project = nn.Linear(cond_dim, output_dim)
        elif projection == "mlp":
            self.project = nn.Sequential(nn.Linear(cond_dim, output_dim), nn.SiLU(), nn.Linear(output_dim, output_dim))
        else:
            self.project = nn.Identity()
        self.uncond_vector = nn.Parameter(torch.zeros(output_dim)) if uncond_type == "learned" else None

    def apply_cond(self, *inputs: Any) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, inputs: tuple | None) -> torch.Tensor:
        if inputs is None:
            assert self.uncond_vector is not None, f"conditioner '{self.name}' has no unconditional vector"
            return self.uncond_vector.data.view(1, 1, -1)
        return self.project(self.apply_cond(*inputs))


class EspeakPhonemeConditioner(Conditioner):
    def __init__(self, output_dim: int, **kwargs):
        super().__init__(output_dim, **kwargs)
        self.phoneme_embedder = nn.Embedding(_N_SPECIAL + len(SYMBOLS), output_dim)

    def apply_cond(self, texts: list[str], languages: list[str]) -> torch.Tensor:
        ids, _ = tokenize_phonemes(phonemize(texts, languages))
        return self.phoneme_embedder(ids.to(self.phoneme_embedder.weight.device))


class FourierConditioner(Conditioner):
    """Random Fourier features of a normalised scalar/vector (zonos/conditioning.py:388-441)."""

    def __init__(self, output_dim: int, input_dim: int = 1, std: float = 1.0, min_val: float = 0.0, max_val: float = 1.0, **kwargs):
        assert output_dim % 2 == 0
        super().__init__(output_dim, **kwargs)
        self.register_buffer("weight", torch.randn(output_dim // 2, input_dim) * std)
        self.input_dim, self.min_val, self.max_val = input_dim, min_val, max_val

    def apply_cond(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-1] == self.input_dim
        x = (x - self.min_val) / (self.max_val - self.min_val)
        f = 2 * torch.pi * x.to(self.weight.dtype) @ self.weight.T
        return torch.cat([f.cos(), f.sin()], dim=-1)


class IntegerConditioner(Conditioner):
    def __init__(self, output_dim: int, min_val: int = 0, max_val: int = 512, **kwargs):
        super().__init__(output_dim, **kwargs)
        self.min_val, self.max_val = min_val, max_val
        self.int_embedder = nn.Embedding(max_val - min_val + 1, output_dim)

    def apply_cond(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-1] == 1
        return self.int_embedder(x.squeeze(-1) - self.min_val)


class PassthroughConditioner(Conditioner):
    def apply_cond(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-1] == self.cond_dim
        return x


_CONDITIONERS = {c.__name__: c for c in (PassthroughConditioner, EspeakPhonemeConditioner, FourierConditioner, IntegerConditioner)}


class PrefixConditioner(Conditioner):
    """Concatenate every conditioner's tokens along the sequence, project, LayerNorm (zonos/conditioning.py:506-522)."""

    def __init__(self, config: PrefixConditionerConfig, output_dim: int):
        super().__init__(output_dim, "prefix", projection=config.projection)
        self.conditioners = nn.ModuleList(_CONDITIONERS[c["type"]](output_dim, **c) for c in config.conditioners)
        self.norm = nn.LayerNorm(output_dim)
        self.required_keys = {c.name for c in self.conditioners if c.uncond_vector is None}

    def forward(self, cond_dict: dict) -> torch.Tensor:
        if not set(cond_dict).issuperset(self.required_keys):
            raise ValueError(f"Missing required keys: {self.required_keys - set(cond_dict)}")
        conds = [c(cond_dict.get(c.name)) for c in self.conditioners]
        bsz = max(len(c) for c in conds)
        assert all(c.shape[0] in (bsz, 1) for c in conds)
        return self.norm(self.project(torch.cat([c.expand(bsz, -1, -1) for c in conds], dim=-2)))


supported_language_codes = [
    'af', 'am', 'an', 'ar', 'as', 'az', 'ba', 'bg', 'bn', 'bpy', 'bs', 'ca', 'cmn', 'cs', 'cy', 'da', 'de', 'el', 'en-029', 'en-gb',
    'en-gb-scotland', 'en-gb-x-gbclan', 'en-gb-x-gbcwmd', 'en-gb-x-rp', 'en-us', 'eo', 'es', 'es-419', 'et', 'eu', 'fa', 'fa-latn',
    'fi', 'fr-be', 'fr-ch', 'fr-fr', 'ga', 'gd', 'gn', 'grc', 'gu', 'hak', 'hi', 'hr', 'ht', 'hu', 'hy', 'hyw', 'ia', 'id', 'is', 'it',
    'ja', 'jbo', 'ka', 'kk', 'kl', 'kn', 'ko', 'kok', 'ku', 'ky', 'la', 'lfn', 'lt', 'lv', 'mi', 'mk', 'ml', 'mr', 'ms', 'mt', 'my',
    'nb', 'nci', 'ne', 'nl', 'om', 'or', 'pa', 'pap', 'pl', 'pt', 'pt-br', 'py', 'quc', 'ro', 'ru', 'ru-lv', 'sd', 'shn', 'si', 'sk',
    'sl', 'sq', 'sr', 'sv', 'sw', 'ta', 'te', 'tn', 'tr', 'tt', 'ur', 'uz', 'vi', 'vi-vn-x-central', 'vi-vn-x-south', 'yue']


def make_cond_dict(text: str = "It would be nice to have time for testing, indeed.", language: str = "en-us", speaker: torch.Tensor | None = None,
                   emotion: list[float] = [0.3077, 0.0256, 0.0256, 0.0256, 0.0256, 0.0256, 0.2564, 0.3077], fmax: float = 22050.0,
                   pitch_std: float = 20.0, speaking_rate: float = 15.0, vqscore_8: list[float] = [0.78] * 8, ctc_loss: float = 0.0,
                   dnsmos_ovrl: float = 4.0, speaker_noised: bool = False, unconditional_keys: Iterable[str] = {"vqscore_8", "dnsmos_ovrl"},
                   device: torch.device | str = "cuda") -> dict:
    """Same keys, shapes ([1,1,n]) and normalisation as zonos/conditioning.py:545-644."""
    lang = language.lower()
    assert lang in supported_language_codes, f"Unsupported language: {language}"
    cond = {"espeak": ([text], [language]), "speaker": speaker, "emotion": emotion, "fmax": fmax, "pitch_std": pitch_std,
            "speaking_rate": speaking_rate, "language_id": supported_language_codes.index(lang), "vqscore_8": vqscore_8,
            "ctc_loss": ctc_loss, "dnsmos_ovrl": dnsmos_ovrl, "speaker_noised": int(speaker_noised)}
    for k in unconditional_keys:
        cond.pop(k, None)
    for k, v in cond.items():
        if isinstance(v, (float, int, list)):
            v = torch.tensor(v)
        if isinstance(v, torch.Tensor):
            cond[k] = v.view(1, 1, -1).to(device)
        if k == "emotion":
            cond[k] = cond[k] / cond[k].sum(dim=-1)
    return cond


# ---- cache (zonos/utilities/conditioning_cache.py:13-193) ------------------------------------------------------------
def _cache_key(cond_dict: dict, uncond_dict: dict | None) -> str:
    h = hashlib.sha512()
    for d in (cond_dict, uncond_dict or {}):
        for k in sorted(d):
            v = d[k]
            h.update(k.encode())
            h.update(v.detach().cpu().numpy().tobytes() if isinstance(v, torch.Tensor) else repr(v).encode())
    return h.hexdigest()


def prepare_conditioning_with_cache(prefix_conditioner: PrefixConditioner, cond_dict: dict, uncond_dict: dict | None = None,
                                    use_cache: bool = False, cfg_scale: float = 1.0, cache: dict | None = None,
                                    max_size: int = 32) -> torch.Tensor:
    """cfg_scale == 1: [B, Lc, D]; otherwise cat([cond, uncond]) on the batch (conditioning_cache.py:160-193)."""
    key = None
    if use_cache and cache is not None:
        key = _cache_key(cond_dict, uncond_dict) + f"|{cfg_scale == 1.0}"
        if key in cache:
            if isinstance(cache, OrderedDict):
                cache.move_to_end(key)
            return cache[key]
    if cfg_scale == 1.0:
        out = prefix_conditioner(cond_dict)
    else:
        if uncond_dict is None:
            uncond_dict = {k: cond_dict[k] for k in prefix_conditioner.required_keys}
        out = torch.cat([prefix_conditioner(cond_dict), prefix_conditioner(uncond_dict)])
    if key is not None:
        cache[key] = out
        while len(cache) > max_size:
            cache.pop(next(iter(cache)))
    return out
