"""The reference's corpus readers (REF/corpus/*.py, REF/data.py:48-78) and collate arithmetic (REF/data.py:11-25).

Same directory conventions, transcript lookups, text normalisation and ORDER (sorted by transcript length, longest first;
TED-LIUM shortest first) as the reference's Dataset classes, so `main.py --dataset_name librispeech --dataset_dir ...`
walks the utterances the way the reference does.  The items are `FileUtterance`s with the interface of
`data.Utterance` (index, n_samples, text, duration, audio()), so the sequential loop and the batched runner take
either.  What the collate function does to the waveform (mono flatten, resample to 16 kHz, clamp to 600 000 samples,
`wav += extra_noise * randn`) is in `read_audio` / `FileUtterance.audio`; normalisation, noise (batched path) and
everything after it run on the device.

Audio decoding is the one third-party piece: `soundfile` if importable, else `torchaudio.load`, else (WAV only) the
standard-library reader below.  The image this repo is built in has neither soundfile nor torchcodec, so FLAC
(LibriSpeech) needs one of them at run time; WAV corpora (CHiME-3, TED-LIUM segments) work everywhere.
"""
from __future__ import annotations

import os
import re
import wave
from pathlib import Path
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .data import MAX_SAMPLES, SAMPLE_RATE


# --------------------------------------------------------------------------------------------------
# audio files
# --------------------------------------------------------------------------------------------------
def _read_wav_stdlib(path: str) -> Tuple[np.ndarray, int]:
    """PCM WAV (8 / 16 / 24 / 32 bit) -> float32 [channels, frames] in [-1, 1), like torchaudio.load's default."""
    with wave.open(path, "rb") as f:
        nch, width, sr, n = f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()
        raw = f.readframes(n)
    if width == 2:
        a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        a = ((v ^ 0x800000) - 0x800000).astype(np.float32) / 8388608.0
    else:
        raise ValueError(f"{path}: unsupported sample width {width}")
    return a.reshape(-1, nch).T.copy(), sr


def _decode(path: str) -> Tuple[np.ndarray, int]:
    try:
        import soundfile as sf
        a, sr = sf.read(path, dtype="float32", always_2d=True)
        return a.T.copy(), sr
    except ImportError:
        pass
    if path.lower().endswith(".wav"):
        try:
            return _read_wav_stdlib(path)
        except wave.Error:
            pass                                   # not plain PCM: let torchaudio try
    try:
        import torchaudio
        w, sr = torchaudio.load(path)
        return w.numpy(), int(sr)
    except ImportError as e:
        raise RuntimeError(f"cannot decode {path}: install `soundfile` (or torchcodec for torchaudio.load); "
                           f"plain PCM .wav files need neither") from e


def read_audio(path: str, max_len: int = MAX_SAMPLES) -> np.ndarray:
    """REF/data.py:15-21 `audio_reader` without the noise: decode, resample to 16 kHz with torchaudio's Resample (the
    reference's), flatten ALL channels into one vector (`wav.reshape(-1)`, as the reference does), clamp to max_len."""
    wav, sr = _decode(str(path))
    if sr != SAMPLE_RATE:
        import torch
        import torchaudio
        wav = torchaudio.transforms.Resample(sr, SAMPLE_RATE)(torch.from_numpy(wav)).numpy()
    wav = np.ascontiguousarray(wav, dtype=np.float32).reshape(-1)
    if wav.shape[-1] >= max_len:
        wav = wav[:max_len]
    return wav


def _n_samples_fast(path: str) -> Optional[int]:
    """Length after resampling and clamping from the file header alone (mono/stereo PCM WAV); None = must decode."""
    if not str(path).lower().endswith(".wav"):
        return None
    try:
        with wave.open(str(path), "rb") as f:
            nch, sr, n = f.getnchannels(), f.getframerate(), f.getnframes()
    except (wave.Error, EOFError):
        return None
    if sr != SAMPLE_RATE:
        return None
    return min(n * nch, MAX_SAMPLES)


class FileUtterance:
    """One (audio file, transcript) pair with the interface of data.Utterance."""

    def __init__(self, index: int, path, text: str, extra_noise: float = 0.0, seed: int = 0):
        self.index, self.path, self.text, self.extra_noise, self.seed = index, str(path), text, extra_noise, seed
        self._n: Optional[int] = None

    @property
    def name(self) -> str:                                   # REF/data.py:36
        return self.path.split('/')[-1].split('.')[0]

    @property
    def n_samples(self) -> int:
        if self._n is None:
            self._n = _n_samples_fast(self.path)
            if self._n is None:
                self._n = len(read_audio(self.path))
        return self._n

    @property
    def duration(self) -> float:
        return self.n_samples / SAMPLE_RATE

    def audio(self, with_noise: bool = True) -> np.ndarray:
        """What the reference's collate function hands to the processor (REF/data.py:15-23).  with_noise=False leaves
        the extra_noise term out (the batched runner adds it on the device)."""
        wav = read_audio(self.path)
        self._n = len(wav)
        if with_noise and self.extra_noise > 0:
            rng = np.random.default_rng(self.seed)
            wav = wav + (self.extra_noise * rng.standard_normal(len(wav))).astype(np.float32)
        return wav


# --------------------------------------------------------------------------------------------------
# corpora
# --------------------------------------------------------------------------------------------------
def _sorted_by_text(files: Sequence, texts: Sequence[str], ascending: bool) -> List[Tuple[str, str]]:
    """REF/corpus/*.py: `sorted(zip(file_list, text), reverse=not ascending, key=lambda x: len(x[1]))` (stable)."""
    return sorted(zip(files, texts), reverse=not ascending, key=lambda x: len(x[1]))


def librispeech_read_text(file: str) -> Optional[str]:
    """REF/corpus/librispeech.py:8-19: the line of <speaker>-<chapter>.trans.txt whose first field is the utterance id."""
    src_file = '-'.join(file.split('-')[:-1]) + '.trans.txt'
    idx = file.split('/')[-1].split('.')[0]
    with open(src_file, 'r') as fp:
        for line in fp:
            if idx == line.split(' ')[0]:
                return line[:-1].split(' ', 1)[1]
    return None


def librispeech(path: str, split=None, ascending: bool = False) -> List[Tuple[str, str]]:
    """REF/corpus/librispeech.py:22-40.  The reference overwrites `split` with ['test-other'] (:28); so does the default."""
    split = ['test-other'] if split is None else split
    file_list = []
    for s in split:
        file_list += list(Path(os.path.join(path, s)).rglob("*.flac"))
    text = [librispeech_read_text(str(f)) for f in file_list]
    return _sorted_by_text(file_list, text, ascending)


CHIME_SPLITS = ['et05_bus_real', 'et05_bus_simu', 'et05_caf_real', 'et05_caf_simu', 'et05_ped_simu', 'et05_str_real',
                'et05_str_simu']                             # REF/corpus/CHiME.py:29 (et05_ped_real is absent there too)


def chime_read_text(tpath: str, file: str) -> Optional[str]:
    """REF/corpus/CHiME.py:9-18: first line of <tpath>/<split>/<name>.trn, without its leading utterance-id field."""
    txt = os.path.join(tpath, "".join("/".join(file.split('/')[-2:]).split(".")[:-1]) + '.trn')
    with open(txt, 'r') as fp:
        for line in fp:
            return ' '.join(line.split(' ')[1:]).strip('\n')
    return None


def chime(path: str, enhance: bool = False, ascending: bool = False) -> List[Tuple[str, str]]:
    """REF/corpus/CHiME.py:22-52 (CHiME-3 et05, 16 kHz 'enhanced' directory; enhance = the se_wav sub-directories)."""
    apath = path + "/data/audio/16kHz/enhanced"
    tpath = path + "/data/transcriptions"
    file_list = []
    for s in CHIME_SPLITS:
        file_list += list(Path(os.path.join(apath, s)).glob("*.wav"))
    text = [chime_read_text(tpath, str(f)) for f in file_list]
    if enhance:
        file_list = []
        for s in CHIME_SPLITS:
            file_list += list(Path(os.path.join(os.path.join(apath, s), 'se_wav')).glob("*.wav"))
    return _sorted_by_text(file_list, text, ascending)


def ted_read_text(tpath: str, file: str) -> Optional[str]:
    """REF/corpus/ted.py:9-19: first line of <tpath>/<name>.txt (every 'wav' in the file name becomes 'txt')."""
    txt = os.path.join(tpath, file.split('/')[-1].replace('wav', 'txt'))
    with open(txt, 'r') as fp:
        for line in fp:
            return line.strip('\n')
    return None


def ted(path: str, enhance: bool = False, ascending: bool = True) -> List[Tuple[str, str]]:
    """REF/corpus/ted.py:23-58 (TED-LIUM 2 test, pre-segmented by REF/preprocess/preprocess_ted.py): SHORTEST first, files
    whose transcript is empty are dropped."""
    apath, tpath = path + "/wav_segment", path + "/transcription"
    file_list = list(Path(os.path.join(apath, 'se_wav') if enhance else apath).glob("*.wav"))
    kept, text = [], []
    for f in file_list:
        t = ted_read_text(tpath, str(f))
        if t is not None:
            kept.append(f)
            text.append(t)
    return _sorted_by_text(kept, text, ascending)


def commonvoice_preprocess_text(text) -> str:
    """REF/corpus/commonvoice.py:12-24."""
    text = str(text)
    for a, b in (("i.e.", "that is"), ("e.g.", "for example"), ("Mr.", "Mister"), ("Mrs.", "Mistress"), ("Dr.", "Doctor"),
                 ("-", " ")):
        text = text.replace(a, b)
    text = text.upper()
    text = re.sub("[^ A-Z']", "", text)
    return ' '.join(text.split())


def commonvoice(path: str, ascending: bool = False) -> List[Tuple[str, str]]:
    """REF/corpus/commonvoice.py:26-44: <path>/test.tsv (columns `path`, `sentence`), audio under <path>/clips."""
    import pandas as pd
    df = pd.read_csv(path + "/test.tsv", sep='\t')
    text = df['sentence'].apply(commonvoice_preprocess_text).values
    file_list = [os.path.join(path + "/clips", f) for f in df['path'].values]
    return _sorted_by_text(file_list, text, ascending)


_CORPORA: dict = {"librispeech": librispeech, "chime": chime, "ted": ted, "commonvoice": commonvoice}


def create_dataset(split, name: str, path: str, batch_size: int = 1, extra_noise: float = 0.0) -> List[FileUtterance]:
    """REF/data.py:48-69 + :72-78 (`load_dataset`): the corpus as a list of utterances in the reference's order.
    `split` is accepted and ignored exactly where the reference ignores it (every Dataset class overwrites it)."""
    fn: Optional[Callable] = _CORPORA.get(name.lower())
    if fn is None:
        raise NotImplementedError(name)                      # REF/data.py:62-63
    pairs = fn(path)
    print(f'[INFO]    There are {len(pairs)} samples.')
    return [FileUtterance(i, f, t, extra_noise=extra_noise, seed=1000 + i) for i, (f, t) in enumerate(pairs)]
