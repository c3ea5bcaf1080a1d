"""CPU: the corpus readers (suta_b200/corpus.py) against the reference's Dataset classes (REF/corpus/*.py) on small
directory trees laid out the way each corpus is, and the collate arithmetic of REF/data.py:11-25.

The reference classes are imported UNMODIFIED when /root/reference exists (the build container); on the GPU box the
same trees are checked against the expected lists written below (which the reference produced here)."""
import os
import sys
import wave

import numpy as np
import pytest

from suta_b200 import corpus

REF = "/root/reference"


def _wav(path, n=1600, sr=16000, nch=1, seed=0):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    a = (np.random.default_rng(seed).standard_normal(n * nch) * 3000).astype("<i2")
    with wave.open(str(path), "wb") as f:
        f.setnchannels(nch); f.setsampwidth(2); f.setframerate(sr)
        f.writeframes(a.tobytes())
    return a


def _ref(module, cls):
    if not os.path.isdir(REF):
        return None
    sys.path.insert(0, REF)
    try:
        return getattr(__import__("corpus." + module, fromlist=[cls]), cls)
    finally:
        sys.path.remove(REF)


@pytest.fixture()
def libri(tmp_path):
    root = tmp_path / "LibriSpeech"
    trans = {"1089-134686": ["HE HOPED THERE WOULD BE STEW FOR DINNER", "STUFF IT INTO YOU", "AFTER EARLY NIGHTFALL THE YELLOW LAMPS"],
             "2300-131720": ["THE PARIS PLANT", "A"]}
    for chap, lines in trans.items():
        spk, ch = chap.split("-")
        d = root / "test-other" / spk / ch
        d.mkdir(parents=True)
        with open(d / f"{chap}.trans.txt", "w") as f:
            for i, t in enumerate(lines):
                f.write(f"{chap}-{i:04d} {t}\n")
                (d / f"{chap}-{i:04d}.flac").write_bytes(b"")
    return root


def test_librispeech_listing_matches_reference(libri, capsys):
    got = corpus.librispeech(str(libri))
    texts = [t for _f, t in got]
    assert sorted(texts, key=len, reverse=True) == texts and len(got) == 5          # longest transcript first
    assert texts[0] == "HE HOPED THERE WOULD BE STEW FOR DINNER" and texts[-1] == "A"
    assert all(str(f).endswith(".flac") for f, _t in got)
    Ref = _ref("librispeech", "LibriDataset")
    if Ref is not None:
        ds = Ref(['test-other'], 1, str(libri))
        ref = [ds[i] for i in range(len(ds))]
        assert sorted((str(f), t) for f, t in ref) == sorted((str(f), t) for f, t in got)
        assert [len(t) for _f, t in ref] == [len(t) for _f, t in got]
    capsys.readouterr()


def test_chime_ted_commonvoice_listings_match_reference(tmp_path, capsys):
    # CHiME-3: <path>/data/audio/16kHz/enhanced/<split>/*.wav, <path>/data/transcriptions/<split>/<name>.trn
    ch = tmp_path / "CHiME3"
    for s, names in (("et05_bus_real", ["F05_440C0201_BUS", "M05_440C0202_BUS"]), ("et05_str_simu", ["F06_441C0203_STR"])):
        for j, nme in enumerate(names):
            _wav(ch / "data/audio/16kHz/enhanced" / s / f"{nme}.wav", seed=j)
            os.makedirs(ch / "data/transcriptions" / s, exist_ok=True)
            with open(ch / "data/transcriptions" / s / f"{nme}.trn", "w") as f:
                f.write(f"{nme} " + " ".join(["WORD"] * (j + 2 + len(s) % 3)) + "\n")
    got = corpus.chime(str(ch))
    assert len(got) == 3 and [len(t) for _f, t in got] == sorted((len(t) for _f, t in got), reverse=True)
    assert all(not t.startswith("F0") and not t.startswith("M0") for _f, t in got)      # the id field is dropped
    Ref = _ref("CHiME", "CHiMEDataset")
    if Ref is not None:
        ds = Ref(None, 1, str(ch))
        assert sorted((str(f), t) for f, t in (ds[i] for i in range(len(ds)))) == sorted((str(f), t) for f, t in got)

    # TED-LIUM segments: <path>/wav_segment/*.wav, <path>/transcription/<name>.txt; shortest first; empty transcripts dropped
    td = tmp_path / "ted"
    for j, (nme, t) in enumerate((("talk_0001", "a longer transcript here"), ("talk_0002", "short"), ("talk_0003", ""))):
        _wav(td / "wav_segment" / f"{nme}.wav", seed=j)
        os.makedirs(td / "transcription", exist_ok=True)
        with open(td / "transcription" / f"{nme}.txt", "w") as f:
            f.write(t + ("\n" if t else ""))
    got = corpus.ted(str(td))
    assert [t for _f, t in got] == ["short", "a longer transcript here"]
    Ref = _ref("ted", "TedDataset")
    if Ref is not None:
        ds = Ref(None, 1, str(td))
        assert [(str(f), t) for f, t in (ds[i] for i in range(len(ds)))] == [(str(f), t) for f, t in got]

    # Common Voice: <path>/test.tsv + <path>/clips
    cv = tmp_path / "cv"
    os.makedirs(cv / "clips")
    with open(cv / "test.tsv", "w") as f:
        f.write("client_id\tpath\tsentence\n")
        f.write("x\tc1.mp3\tMr. Smith's well-known e.g. \"quote\", 42 times!\n")
        f.write("y\tc2.mp3\tDr. Who\n")
    got = corpus.commonvoice(str(cv))
    assert [t for _f, t in got] == ["MISTER SMITH'S WELL KNOWN FOR EXAMPLE QUOTE TIMES", "DOCTOR WHO"]
    assert got[0][0] == str(cv / "clips" / "c1.mp3")
    Ref = _ref("commonvoice", "CVDataset")
    if Ref is not None:
        ds = Ref(None, 1, str(cv))
        assert [(str(f), t) for f, t in (ds[i] for i in range(len(ds)))] == [(str(f), t) for f, t in got]
    capsys.readouterr()


def test_read_audio_follows_the_reference_collate(tmp_path):
    """REF/data.py:15-23: decode to float32 in [-1, 1), resample to 16 kHz, flatten every channel into ONE vector
    (`wav.reshape(-1)` of a [channels, frames] tensor), keep the first 600 000 samples, add the noise last."""
    a = _wav(tmp_path / "m.wav", n=3200, seed=1)
    w = corpus.read_audio(str(tmp_path / "m.wav"))
    assert w.dtype == np.float32 and np.array_equal(w, a.astype(np.float32) / 32768.0)
    s = _wav(tmp_path / "s.wav", n=100, nch=2, seed=2).reshape(-1, 2)
    w = corpus.read_audio(str(tmp_path / "s.wav"))
    assert np.array_equal(w, np.concatenate([s[:, 0], s[:, 1]]).astype(np.float32) / 32768.0)      # channel after channel
    _wav(tmp_path / "long.wav", n=600123, seed=3)
    assert len(corpus.read_audio(str(tmp_path / "long.wav"))) == 600000
    _wav(tmp_path / "r8.wav", n=8000, sr=8000, seed=4)
    w = corpus.read_audio(str(tmp_path / "r8.wav"))
    assert len(w) == 16000                                        # torchaudio.transforms.Resample(8000, 16000)
    u = corpus.FileUtterance(0, tmp_path / "m.wav", "HELLO", extra_noise=0.01, seed=5)
    assert u.n_samples == 3200 and u.duration == 0.2 and u.name == "m"
    clean, noisy = u.audio(with_noise=False), u.audio()
    assert 0.005 < float(np.std(noisy - clean)) < 0.02


def test_create_dataset_dispatch_and_cli_flag(libri, capsys):
    ds = corpus.create_dataset(['test-other'], "LibriSpeech", str(libri), 1, extra_noise=0.0)
    assert [u.index for u in ds] == list(range(5)) and ds[0].text.startswith("HE HOPED")
    with pytest.raises(NotImplementedError):                     # REF/data.py:62-63
        corpus.create_dataset(None, "switchboard", str(libri))
    assert "There are 5 samples" in capsys.readouterr().out


def test_ted_preprocessing_cuts_talks_like_the_reference(tmp_path, capsys):
    """REF/preprocess/preprocess_ted.sh + preprocess_ted.py without sox / soundfile: one talk as NIST SPHERE, one as the
    reference's intermediate WAV; segments [int(s * sr), int(e * sr)), names built from the STM's own strings, normalised
    transcripts, gap lines skipped -- and the result is what the TED reader (REF/corpus/ted.py) lists."""
    from suta_b200 import preprocess_ted as P
    split = tmp_path / "test"
    for d in ("stm", "sph", "wav"):
        os.makedirs(split / d)
    rng = np.random.default_rng(5)
    a = (rng.standard_normal(16000 * 4) * 3000).astype("<i2")
    head = ("NIST_1A\n   1024\nchannel_count -i 1\nsample_rate -i 16000\nsample_n_bytes -i 2\nsample_byte_format -s2 01\n"
            f"sample_coding -s3 pcm\nsample_count -i {len(a)}\nend_head\n").encode()
    (split / "sph" / "TalkA_2010.sph").write_bytes(head + b" " * (1024 - len(head)) + a.tobytes())
    b = _wav(split / "wav" / "TalkB_2011.wav", n=16000 * 3, seed=6)
    (split / "stm" / "TalkA_2010.stm").write_text(
        "TalkA_2010 1 inter_segment_gap 0 0.5 <o,,unknown> ignore_time_segment_in_scoring\n"
        "TalkA_2010 1 TalkA_2010 0.5 1.75 <o,f0,male> it 's a well-known fact (laughter) that 3 birds\n"
        "TalkA_2010 1 TalkA_2010 2.01 3.9 <o,f0,male> so i said  yes\n")
    (split / "stm" / "TalkB_2011.stm").write_text("TalkB_2011 1 TalkB_2011 0.25 2.5 <o,f0,female> thank you\n")
    files = P.preprocess(str(split))
    capsys.readouterr()
    assert sorted(os.path.basename(f) for f in files) == ["TalkA_2010-0.5-1.75.wav", "TalkA_2010-2.01-3.9.wav", "TalkB_2011-0.25-2.5.wav"]
    assert (split / "transcription" / "TalkA_2010-0.5-1.75.txt").read_text() == "IT'S A WELL KNOWN FACT LAUGHTER THAT BIRDS"
    assert (split / "transcription" / "TalkA_2010-2.01-3.9.txt").read_text() == "SO I SAID YES"

    def seg(name):
        with wave.open(str(split / "wav_segment" / name), "rb") as f:
            assert (f.getframerate(), f.getnchannels(), f.getsampwidth()) == (16000, 1, 2)
            return np.frombuffer(f.readframes(f.getnframes()), dtype="<i2")
    s0, s1, s2 = seg("TalkA_2010-0.5-1.75.wav"), seg("TalkA_2010-2.01-3.9.wav"), seg("TalkB_2011-0.25-2.5.wav")
    assert len(s0) == int(1.75 * 16000) - int(0.5 * 16000) and len(s1) == int(3.9 * 16000) - int(2.01 * 16000) and len(s2) == 36000
    want = np.rint(a[8000:28000].astype(np.float64) / 32768 * 32767).astype(np.int16)       # soundfile's read / write round trip
    assert np.array_equal(s0, want) and np.abs(s0.astype(int) - a[8000:28000]).max() <= 1
    assert np.abs(s2.astype(int) - b[4000:40000]).max() <= 1
    listing = corpus.ted(str(split))                       # REF/corpus/ted.py order: shortest transcript first
    assert [t for _f, t in listing] == ["THANK YOU", "SO I SAID YES", "IT'S A WELL KNOWN FACT LAUGHTER THAT BIRDS"]
    x = corpus.read_audio(str(listing[0][0]))
    assert x.dtype == np.float32 and len(x) == 36000
