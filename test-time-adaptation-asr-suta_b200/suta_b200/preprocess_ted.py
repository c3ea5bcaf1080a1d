"""TED-LIUM 2 data preparation: what REF/preprocess/preprocess_ted.sh + preprocess_ted.py do, without sox / soundfile.

The reference converts every talk `<split>/sph/<talk>.sph` to WAV with sox (preprocess_ted.sh), then cuts it at the
segment boundaries of `<split>/stm/<talk>.stm` into `<split>/wav_segment/<talk>-<start>-<end>.wav` and writes the
normalised transcript to `<split>/transcription/<talk>-<start>-<end>.txt` (preprocess_ted.py) -- the layout
`suta_b200.corpus.ted` (REF/corpus/ted.py) reads.  Here both steps are one pass over the STM files:

    python -m suta_b200.preprocess_ted /data/TEDLIUM_release2/test

* audio: `<split>/wav/<talk>.wav` if it exists (the reference's intermediate), else `<split>/sph/<talk>.sph` read
  directly (NIST SPHERE with 16-bit PCM samples, the format of the TED-LIUM releases; `shorten`-compressed files need
  sox, as in the reference);
* a segment is samples `[int(start * sr), int(end * sr))` (REF/preprocess/preprocess_ted.py:42-44), written as 16-bit PCM at
  16 kHz; the reference's round trip through float (soundfile reads int16 as x / 32768 and libsndfile writes
  lrint(x * 32767)) is reproduced, so the files carry the same sample values;
* text: upper-case, " '" -> "'", "-" -> " ", everything outside [ A-Z'] dropped, blanks collapsed (:13-20);
  `inter_segment_gap` lines are skipped (:22,33-34).
"""
from __future__ import annotations

import glob
import os
import re
import sys
import wave
from typing import Iterator, List, Tuple

import numpy as np

SAMPLE_RATE = 16000
SKIP = "inter_segment_gap"


def preprocess_text(text: str) -> str:
    """REF/preprocess/preprocess_ted.py:13-20."""
    text = text.upper()
    text = text.replace(" '", "'")
    text = text.replace("-", " ")
    text = re.sub("[^ A-Z']", "", text)
    return " ".join(text.split())


def read_sphere(path: str) -> Tuple[np.ndarray, int]:
    """int16 samples (first channel) and sample rate of a NIST SPHERE file with PCM samples."""
    with open(path, "rb") as f:
        head = f.read(16)
        if not head.startswith(b"NIST_1A"):
            raise ValueError(f"{path}: not a NIST SPHERE file")
        hsize = int(head.split()[1])
        f.seek(0)
        fields = {}
        for line in f.read(hsize).decode("latin-1").splitlines()[2:]:
            parts = line.split(None, 2)
            if parts and parts[0] == "end_head":
                break
            if len(parts) == 3:
                fields[parts[0]] = parts[2]
        coding = fields.get("sample_coding", "pcm")
        if coding != "pcm" or int(fields.get("sample_n_bytes", 2)) != 2:
            raise ValueError(f"{path}: sample_coding {coding!r} / {fields.get('sample_n_bytes')} bytes -- convert it with sox "
                             "(REF/preprocess/preprocess_ted.sh) and put the result under <split>/wav/")
        data = f.read()
    order = "<" if fields.get("sample_byte_format", "01") == "01" else ">"
    x = np.frombuffer(data[:len(data) // 2 * 2], dtype=order + "i2")
    ch = int(fields.get("channel_count", 1))
    if ch > 1:
        x = x[:len(x) // ch * ch].reshape(-1, ch)[:, 0]
    n = int(fields.get("sample_count", len(x)))
    return x[:n].astype(np.int16), int(fields.get("sample_rate", SAMPLE_RATE))


def read_wav16(path: str) -> Tuple[np.ndarray, int]:
    with wave.open(path, "rb") as f:
        if f.getsampwidth() != 2:
            raise ValueError(f"{path}: expected 16-bit PCM")
        x = np.frombuffer(f.readframes(f.getnframes()), dtype="<i2")
        if f.getnchannels() > 1:
            x = x.reshape(-1, f.getnchannels())[:, 0]
        return x.astype(np.int16), f.getframerate()


def _through_float(x: np.ndarray) -> np.ndarray:
    """sf.read (int16 -> x / 32768) followed by sf.write(..., PCM_16) (lrint(x * 32767)): what the reference's files hold."""
    return np.rint(x.astype(np.float64) / 32768.0 * 32767.0).astype(np.int16)


def stm_segments(stm_file: str) -> Iterator[Tuple[str, str, str, str]]:
    """(talk, start, end, normalised text) per STM line, start / end as the strings of the file (they name the outputs)."""
    with open(stm_file, "r") as f:
        for line in f:
            l = line.split()
            if len(l) < 6 or l[2] == SKIP:
                continue
            yield l[0], l[3], l[4], preprocess_text(" ".join(l[6:]))


def preprocess(split_dir: str, verbose: bool = True) -> List[str]:
    """Cut every talk of `<split_dir>/stm/*.stm`; returns the segment files written."""
    out_a, out_t = os.path.join(split_dir, "wav_segment"), os.path.join(split_dir, "transcription")
    os.makedirs(out_a, exist_ok=True)
    os.makedirs(out_t, exist_ok=True)
    written, cur, wav, sr = [], None, None, SAMPLE_RATE
    for stm in sorted(glob.glob(os.path.join(split_dir, "stm", "*.stm"))):
        for talk, s, e, text in stm_segments(stm):
            if talk != cur:
                w = os.path.join(split_dir, "wav", talk + ".wav")
                wav, sr = read_wav16(w) if os.path.exists(w) else read_sphere(os.path.join(split_dir, "sph", talk + ".sph"))
                cur = talk
                if verbose:
                    print(f"{talk}: {len(wav)} samples at {sr} Hz")
            seg = _through_float(wav[int(float(s) * sr):int(float(e) * sr)])
            stem = "-".join([talk, s, e])
            with wave.open(os.path.join(out_a, stem + ".wav"), "wb") as f:
                f.setnchannels(1)
                f.setsampwidth(2)
                f.setframerate(SAMPLE_RATE)              # the reference writes SAMPLE_RATE whatever the talk's rate (:52)
                f.writeframes(seg.astype("<i2").tobytes())
            with open(os.path.join(out_t, stem + ".txt"), "w") as f:
                f.write(text)
            written.append(os.path.join(out_a, stem + ".wav"))
    return written


if __name__ == "__main__":
    if len(sys.argv) != 2:
        print(__doc__)
        sys.exit(2)
    print(f"{len(preprocess(sys.argv[1]))} segments written")
