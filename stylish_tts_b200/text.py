"""Text front of the text -> wav path: the reference's symbol table / ``TextCleaner``
(src/stylish_tts/lib/text_utils.py:8-42, symbols from model.yml:80-84), the phoneme strings of its
``sample_dataset`` (BASELINE.json configs[0]) and the voicepack style lookup of the inference CLI
(src/stylish_tts/tts/cli.py:36-81).  Host-side, tiny, no kernels: it only produces the int64 token tensors and
the three (1,64) style rows that ``Synthesizer`` consumes."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


class TextCleaner:
    """phoneme string -> token ids, ``$`` pad added at both ends (text_utils.py:21-29); unknown characters are
    dropped like the reference does (it logs and continues)."""

    def __init__(self, symbols):
        table = [symbols.pad] + list(symbols.punctuation) + list(symbols.letters) + list(symbols.letters_ipa)
        self.index = {}
        for i, ch in enumerate(table):
            self.index[ch] = i  # later duplicates win, as in the reference's dict construction
        self.pad = symbols.pad

    def __call__(self, text: str) -> List[int]:
        return [self.index[ch] for ch in self.pad + text + self.pad if ch in self.index]


# the 9 + 1 real lines of sample_dataset/{training,validation}-list.txt (phoneme field); the rest is filler text
SAMPLE_PHONEMES: Tuple[str, ...] = (
    "ɔnðə kˈɑːntɹɛɹi",
    "fɚðə fˈɜːst tˈaɪm",
    "æz tˈaɪm pˈæst",
    "ðɪ ˈɜːli jˈɪɹz",
    "hˈɑːɹdli ˈɛnɪwˌʌn",
    "wˌɛn ðæt tˈaɪm ɚɹˈaɪvz",
    "wˈʌn nˈaɪt hiː wʌz mˈɪsɪŋ",
    "ðɛɹ ɪz nˈoʊ mˈædʒɪk fˈɔːɹmjʊlə",
    "wɪð ðˌɛm wɜː jˈʌŋɡɚ mˈɛn ænd wˈɪmɪn",
    "ˈiːtʃ əv ˌaʊɚ stˈɛps hɐz ɐ dˈɛfɪnət ɹᵻlˈeɪʃənʃˌɪp tʊ ˈɛvɹi ˈʌðɚ stˈɛp",
)


def pad_batch(token_lists: Sequence[Sequence[int]]) -> Tuple[torch.Tensor, torch.Tensor]:
    """ragged token lists -> (texts (B,Tmax) int64 padded with id 0, lengths (B,))"""
    lengths = torch.tensor([len(t) for t in token_lists], dtype=torch.int64)
    texts = torch.zeros((len(token_lists), int(lengths.max())), dtype=torch.int64)
    for i, t in enumerate(token_lists):
        texts[i, :len(t)] = torch.tensor(list(t), dtype=torch.int64)
    return texts, lengths


def split_voicepack(voicepack: torch.Tensor):
    """(N, >=192) -> speech / pitch-energy / duration packs of 64 columns each (+ the SBERT columns of a dynamic
    pack, cli.py:45-54)"""
    return voicepack[:, :64], voicepack[:, 64:128], voicepack[:, 128:192], voicepack[:, 192:]


def static_voice_index(n_tokens: int) -> int:
    """cli.py:73 literally: ``max(511, min(2, len(tokens)))`` — which is 511 for every length (the arguments of
    min / max are swapped in the reference); kept as is so that the same style row is selected"""
    return max(511, min(2, n_tokens))


def static_styles(voicepack: torch.Tensor, n_tokens: int):
    """static voicepack (512 rows, one per utterance length): the three (1,64) style rows (cli.py:72-77)"""
    sp, pe, du, _ = split_voicepack(voicepack)
    i = static_voice_index(n_tokens)
    return sp[i:i + 1], pe[i:i + 1], du[i:i + 1]


def dynamic_styles(voicepack: torch.Tensor, indices: torch.Tensor, distances: torch.Tensor):
    """dynamic voicepack: the k nearest rows (indices (1,k), distances (1,k) from the caller's sentence-embedding
    search): inverse-distance weighted speech style, plain means for the other two (cli.py:63-71)"""
    sp, pe, du, _ = split_voicepack(voicepack)
    w = 1.0 / distances
    w = (w / w.sum(dim=1, keepdim=True)).unsqueeze(2).to(sp.dtype)
    return (sp[indices] * w).sum(dim=1), pe[indices].mean(dim=1), du[indices].mean(dim=1)
