"""Helper of test_gpu_e2e.py::test_cuda_graph_chains_give_the_same_bits: adapts a few small batches and dumps what came out.
Run twice by the test (with and without SUTA_NO_GRAPH=1: the switch is read once per process)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "test-time-adaptation-asr-suta_b200"), os.path.join(ROOT, "tests")]

import e2e_checks as E  # noqa: E402
from oracle import suta_oracle as O  # noqa: E402
from suta_b200 import AdaptHyper, SutaEngine  # noqa: E402
from suta_b200.runner import adapt_batch  # noqa: E402
from suta_b200.text import CTCVocab  # noqa: E402
import torch  # noqa: E402


def main(out_path):
    out = {}
    for tag, cfg_name, kw, lens in (("ln", "tiny", {}, [9000, 4000, 12345]), ("feat", "tiny", dict(train_feature=True), [9000, 5000]),
                                    ("lv60", "tiny_lv60", {}, [7000, 9000]), ("sdpl", "tiny", dict(pseudo_label=True), [9000]),
                                    ("all", "tiny", dict(train_all=True), [12000]), ("lv60all", "tiny_lv60", dict(train_all=True), [9000])):
        ocfg, mcfg = E._cfgs(cfg_name)
        sd = O.init_weights(ocfg, 3, blank_bias=0.5, ln_jitter=0.1, **({"special_bias": -10.0} if tag == "sdpl" else {}))
        mult = E._mult(ocfg, kw.get("train_feature", False), False, kw.get("train_all", False))
        eng = SutaEngine(mcfg, sd, trainable_mult=mult, **kw)
        hp = AdaptHyper(pl_coef=1.0 if tag == "sdpl" else 0.0)
        for rep in range(2):                                  # two batches through one engine: the chains are re-recorded
            wavs = [O.synth_audio(n + 100 * rep, 40 + i) for i, n in enumerate(lens)]
            eng.begin_batch(wavs)
            host = torch.zeros(eng.total_samples, dtype=torch.float32)
            for w, o in zip(wavs, eng.sample_off):
                host[o:o + len(w)] = torch.from_numpy(w)
            texts = adapt_batch(eng, host.pin_memory(), eng.lengths, 6, hp, CTCVocab())
            out[f"{tag}{rep}_logits"] = eng.logits().cpu().numpy().copy()
            out[f"{tag}{rep}_params"] = eng.params().cpu().numpy().copy()
            out[f"{tag}{rep}_texts"] = np.asarray([" | ".join(texts[k]) for k in sorted(texts)])
        out[f"{tag}_replays"] = np.asarray(eng.graph_replays)
        out[f"{tag}_launches"] = np.asarray(eng.launch_count)
        eng.close()
    np.savez(out_path, **out)


if __name__ == "__main__":
    main(sys.argv[1])
