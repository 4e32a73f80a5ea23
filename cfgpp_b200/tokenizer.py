"""Host-side prompt tokenisation for the CLIP text towers (the `self.tokenizer(...)` / `tokenizer_1` / `tokenizer_2`
calls of the reference: latent_diffusion.py:101-113, latent_sdxl.py:78-84 — transformers `CLIPTokenizer` with
`padding='max_length', max_length=77, truncation=True`).

`ClipBPETokenizer` restates the published CLIP byte-pair-encoding scheme (lower-cased, whitespace-normalised text;
the CLIP word pattern; byte -> printable-unicode alphabet; ranked merges with an end-of-word marker) on the standard
`vocab.json` / `merges.txt` pair of a CLIP checkpoint. tests/test_text_encoder_cpu.py checks it token for token against
transformers' own `CLIPTokenizer` on a generated vocabulary. No vocabulary file exists offline, so the default solver
path uses `HashTokenizer`: a deterministic word -> id stand-in with the same framing (<|startoftext|>, at most 75
tokens, <|endoftext|>, padding) — it exercises the encoder with well-formed ids but is NOT CLIP's vocabulary.
"""
from __future__ import annotations

import hashlib
import html
import json
import unicodedata
from functools import lru_cache
from typing import Dict, List, Sequence, Tuple

import regex

BOS_TOKEN = "<|startoftext|>"
EOS_TOKEN = "<|endoftext|>"
_WORD_PATTERN = regex.compile(
    r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    """The reversible byte -> printable character table of byte-level BPE: printable Latin-1 bytes map to themselves,
    the remaining ones to code points from 256 upwards, in byte order."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


def clean_text(text: str) -> str:
    text = html.unescape(html.unescape(text))
    text = unicodedata.normalize("NFC", text)
    return regex.sub(r"\s+", " ", text).strip().lower()


class _Framing:
    """Shared `padding='max_length', truncation=True` framing."""
    model_max_length = 77
    bos_token_id: int
    eos_token_id: int
    pad_token_id: int

    def _frame(self, ids: Sequence[int]) -> List[int]:
        body = list(ids)[: self.model_max_length - 2]
        out = [self.bos_token_id] + body + [self.eos_token_id]
        return out + [self.pad_token_id] * (self.model_max_length - len(out))

    def __call__(self, prompts) -> List[List[int]]:
        if isinstance(prompts, str):
            prompts = [prompts]
        return [self._frame(self.tokenize(p)) for p in prompts]

    def pooled_index(self, rows: Sequence[Sequence[int]], eos_rule_id: int = 2) -> List[int]:
        """Row position transformers' CLIPTextTransformer pools at: legacy configs (`eos_token_id == 2`, which the SD /
        SDXL text-encoder configs still carry) take the arg-max id, newer ones the first `eos_token_id`."""
        if eos_rule_id == 2:
            return [max(range(len(r)), key=lambda i: (r[i], -i)) for r in rows]
        return [list(r).index(eos_rule_id) if eos_rule_id in r else 0 for r in rows]


class ClipBPETokenizer(_Framing):
    def __init__(self, vocab_file: str, merges_file: str, pad_token: str = EOS_TOKEN):
        with open(vocab_file, encoding="utf-8") as f:
            self.encoder: Dict[str, int] = json.load(f)
        with open(merges_file, encoding="utf-8") as f:
            lines = f.read().strip().split("\n")
        if lines and lines[0].startswith("#"):
            lines = lines[1:]
        merges = [tuple(l.split()) for l in lines if l.strip()]
        self.ranks: Dict[Tuple[str, str], int] = {m: i for i, m in enumerate(merges)}
        self.byte_encoder = bytes_to_unicode()
        self.bos_token_id = self.encoder[BOS_TOKEN]
        self.eos_token_id = self.encoder[EOS_TOKEN]
        self.pad_token_id = self.encoder[pad_token]
        self.unk_token_id = self.eos_token_id
        self._cache: Dict[str, Tuple[str, ...]] = {BOS_TOKEN: (BOS_TOKEN,), EOS_TOKEN: (EOS_TOKEN,)}

    @property
    def vocab_size(self) -> int:
        return len(self.encoder)

    def _bpe(self, token: str) -> Tuple[str, ...]:
        if token in self._cache:
            return self._cache[token]
        word: List[str] = list(token[:-1]) + [token[-1] + "</w>"]
        while len(word) > 1:
            best, best_rank = None, None
            for pair in zip(word[:-1], word[1:]):
                r = self.ranks.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = pair, r
            if best is None:
                break
            merged, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and (word[i], word[i + 1]) == best:
                    merged.append(word[i] + word[i + 1])
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = merged
        out = tuple(word)
        self._cache[token] = out
        return out

    def tokenize(self, text: str) -> List[int]:
        ids: List[int] = []
        for tok in _WORD_PATTERN.findall(clean_text(text)):
            sym = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder.get(piece, self.unk_token_id) for piece in self._bpe(sym))
        return ids


class HashTokenizer(_Framing):
    """Offline stand-in: the CLIP word pattern splits the prompt, every piece hashes to an id in [1, vocab - 3].
    Framing, padding and special ids follow the real tokenizers (CLIP-L pads with <|endoftext|> = 49407, the SDXL
    tokenizer_2 with id 0)."""

    def __init__(self, vocab_size: int = 49408, pad_token_id: int = 49407):
        self.vocab_size = vocab_size
        self.bos_token_id = vocab_size - 2
        self.eos_token_id = vocab_size - 1
        self.pad_token_id = pad_token_id

    def tokenize(self, text: str) -> List[int]:
        out = []
        for tok in _WORD_PATTERN.findall(clean_text(text)):
            h = int.from_bytes(hashlib.sha256(tok.encode("utf-8")).digest()[:8], "little")
            out.append(1 + h % (self.vocab_size - 3))
        return out
