"""CLIP's byte-pair-encoding tokenizer as a fluxion module (first child of every `CLIPTextEncoder`).

Contract - constructor arguments, ``forward`` / ``tokenize_str`` / ``encode``, token ids, padding - from
/root/reference/src/refiners/foundationals/clip/tokenizer.py:12-129.  Same ids as OpenAI's tokenizer for ASCII text (and as the reference's for any text): text is
lower-cased, whitespace collapsed, split by the (non-unicode) token pattern, every token's UTF-8 bytes mapped to printable
characters, merged by rank, and framed by <|startoftext|> ... <|endoftext|>, then padded to ``sequence_length``.

The merge table (``bpe_simple_vocab_16e6.txt.gz``, 48 894 ranked pairs) is DATA published with OpenAI CLIP and shipped inside
the reference package; it is not part of this repository.  It is looked for at ``vocabulary_path``, then ``$RB200_CLIP_VOCAB``,
then next to this file, then in an installed ``refiners`` package, and only when text is first tokenized - building an
encoder (for its weights, or to feed it token ids directly: ``forward`` passes integer tensors through) needs no file.
"""

from __future__ import annotations

import gzip
import os
import re
from pathlib import Path
from typing import Iterator

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl

VOCABULARY_FILE = "bpe_simple_vocab_16e6.txt.gz"
MERGES = 49152 - 256 - 2  # ranked pairs that make up the 49 408-entry vocabulary with the 512 byte tokens and 2 specials


def _byte_alphabet() -> dict[int, str]:
    """One character per byte value, in VOCABULARY order: the visible Latin-1 ranges first, then the other 68 bytes.  Every
    byte stands for the code point of its own value - the reference's table (tokenizer.py:71-81); OpenAI's original moves
    the 68 invisible bytes to code points 256+, which gives other ids for text containing e.g. UTF-8 continuation bytes
    0x80-0xA0 - and the order fixes the ids of the 512 single-byte tokens."""
    visible = [*range(ord("!"), ord("~") + 1), *range(ord("¡"), ord("¬") + 1), *range(ord("®"), ord("ÿ") + 1)]
    hidden = [b for b in range(256) if b not in set(visible)]
    return {b: chr(b) for b in [*visible, *hidden]}


class CLIPTokenizer(fl.Module):
    def __init__(
        self,
        vocabulary_path: str | Path | None = None,
        sequence_length: int = 77,
        start_of_text_token_id: int = 49406,
        end_of_text_token_id: int = 49407,
        pad_token_id: int = 49407,
    ) -> None:
        super().__init__()
        self.vocabulary_path = vocabulary_path
        self.sequence_length = sequence_length
        self.start_of_text_token_id = start_of_text_token_id
        self.end_of_text_token_id = end_of_text_token_id
        self.pad_token_id = pad_token_id
        self.byte_to_unicode_mapping = _byte_alphabet()
        self.byte_decoder = {v: k for k, v in self.byte_to_unicode_mapping.items()}
        self.token_pattern = re.compile(
            r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[a-zA-Z]+|[0-9]|(?:[^\s\w]|_)+""", flags=re.IGNORECASE
        )
        self._tables: tuple[dict[str, int], dict[tuple[str, str], int]] | None = None
        self.byte_pair_encoding_cache: dict[str, str] = {"": ""}

    # -- the merge table -------------------------------------------------------------------------------------
    def _locate(self) -> Path:
        candidates = [self.vocabulary_path, os.environ.get("RB200_CLIP_VOCAB"), Path(__file__).resolve().parent / VOCABULARY_FILE]
        try:
            import refiners.foundationals.clip as installed  # type: ignore[import-not-found]

            candidates.append(Path(installed.__file__).resolve().parent / VOCABULARY_FILE)
        except Exception:
            pass
        for candidate in candidates:
            if candidate and Path(candidate).is_file():
                return Path(candidate)
        raise FileNotFoundError(
            f"CLIPTokenizer needs the BPE merge table {VOCABULARY_FILE} (published with OpenAI CLIP): pass vocabulary_path=..., "
            "set RB200_CLIP_VOCAB, or feed the text encoder token ids"
        )

    def _load(self) -> tuple[dict[str, int], dict[tuple[str, str], int]]:
        if self._tables is None:
            lines = gzip.open(self._locate()).read().decode("utf-8").split("\n")
            merges = [tuple(line.split()) for line in lines[1 : MERGES + 1]]
            letters = [*self.byte_to_unicode_mapping.values()]
            vocabulary = [*letters, *(c + "</w>" for c in letters), *("".join(m) for m in merges), "", ""]
            self._tables = ({tok: i for i, tok in enumerate(vocabulary)}, {m: rank for rank, m in enumerate(merges)})  # type: ignore[misc]
        return self._tables

    @property
    def token_to_id_mapping(self) -> dict[str, int]:
        return self._load()[0]

    @property
    def byte_pair_encoding_ranks(self) -> dict[tuple[str, str], int]:
        return self._load()[1]

    # -- text -> ids -----------------------------------------------------------------------------------------
    def forward(self, text: str | list[str] | Tensor) -> Tensor:
        if isinstance(text, Tensor):  # already token ids
            assert not torch.is_floating_point(text) and text.ndim == 2, "token ids must be an integer [batch, length] tensor"
            return text
        if isinstance(text, str):
            return self.tokenize_str(text)
        assert isinstance(text, list), f"Expected type `str` or `list[str]`, got {type(text)}"
        return torch.cat([self.tokenize_str(one) for one in text])

    def tokenize_str(self, text: str) -> Tensor:
        ids = self.encode(text, max_length=self.sequence_length)
        assert len(ids) <= self.sequence_length, f"Text is too long ({len(text)}): {len(ids)} tokens > {self.sequence_length}"
        padded = torch.full((1, self.sequence_length), self.pad_token_id, dtype=ids.dtype)
        padded[0, : len(ids)] = ids
        return padded

    def byte_pair_encoding(self, token: str) -> str:
        """The sub-words of one pattern match (already in the byte alphabet), space separated: start from characters (the
        last one carrying the end-of-word mark) and keep fusing the adjacent pair of lowest rank until none is ranked."""
        done = self.byte_pair_encoding_cache.get(token)
        if done is not None:
            return done
        ranks = self.byte_pair_encoding_ranks
        parts = [*token[:-1], token[-1] + "</w>"]
        while len(parts) > 1:
            rank, at = min((ranks.get((a, b), MERGES), i) for i, (a, b) in enumerate(zip(parts, parts[1:])))
            if rank == MERGES:
                break
            parts[at : at + 2] = [parts[at] + parts[at + 1]]
        self.byte_pair_encoding_cache[token] = " ".join(parts)
        return self.byte_pair_encoding_cache[token]

    def _ids(self, text: str) -> Iterator[int]:
        ids = self.token_to_id_mapping
        for match in re.findall(self.token_pattern, re.sub(r"\s+", " ", text.lower())):
            spelled = "".join(self.byte_to_unicode_mapping[b] for b in match.encode("utf-8"))
            for piece in self.byte_pair_encoding(spelled).split(" "):
                yield ids[piece]

    def encode(self, text: str, max_length: int | None = None) -> Tensor:
        """<start> ids <end>; with ``max_length`` the ids are cut so that the frame still fits."""
        body = [*self._ids(text)]
        if max_length:
            assert max_length >= 2
            body = body[: max_length - 2]
        return torch.tensor([self.start_of_text_token_id, *body, self.end_of_text_token_id])
