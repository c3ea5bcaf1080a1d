"""CTC vocabulary and id -> text mapping (REF/vocab.json; HF/tokenization_wav2vec2.py:296-358).

The collapse (repeat removal + blank drop) happens on the device (csrc/decode.cu); this is the host half:
word delimiter -> space, join, strip; special tokens 1,2,3 are printed literally because the reference calls
batch_decode with skip_special_tokens=False (REF/main.py:334)."""
from __future__ import annotations

import json
from typing import List, Optional, Sequence

DEFAULT_VOCAB = {"<pad>": 0, "<s>": 1, "</s>": 2, "<unk>": 3, "|": 4, "E": 5, "T": 6, "A": 7, "O": 8, "N": 9, "I": 10,
                 "H": 11, "S": 12, "R": 13, "D": 14, "L": 15, "U": 16, "M": 17, "W": 18, "C": 19, "F": 20, "G": 21,
                 "Y": 22, "P": 23, "B": 24, "V": 25, "K": 26, "'": 27, "X": 28, "J": 29, "Q": 30, "Z": 31}


class CTCVocab:
    def __init__(self, vocab: Optional[dict] = None, blank: str = "<pad>", delimiter: str = "|"):
        vocab = dict(vocab or DEFAULT_VOCAB)
        self.id_to_tok = [None] * len(vocab)
        for t, i in vocab.items():
            self.id_to_tok[i] = t
        self.blank_id = vocab[blank]
        self.delim_id = vocab[delimiter]
        if self.blank_id != 0:
            raise ValueError("the device decoder assumes the CTC blank is id 0 (REF/main.py:178)")

    @staticmethod
    def from_json(path: str) -> "CTCVocab":
        with open(path) as f:
            return CTCVocab(json.load(f))

    def ids_to_text(self, collapsed: Sequence[int]) -> str:
        return "".join(" " if i == self.delim_id else self.id_to_tok[i] for i in collapsed).strip()

    def batch_to_text(self, batch: Sequence[Sequence[int]]) -> List[str]:
        return [self.ids_to_text(x) for x in batch]
