"""Run every operator check and an end-to-end comparison on the GPU box; print metrics (never asserts)."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "test-time-adaptation-asr-suta_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402


def main():
    import gpu_checks as K
    only = sys.argv[1:]
    print("device:", torch.cuda.get_device_name(0), flush=True)
    import e2e_checks as E
    for name, fn in K.ALL + E.ALL:
        if only and not any(o in name for o in only):
            continue
        t0 = time.time()
        try:
            r = fn()
            print(f"[{name}] {json.dumps(r)}  ({time.time() - t0:.2f}s)", flush=True)
        except Exception:
            print(f"[{name}] EXCEPTION\n{traceback.format_exc()}", flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print("CUDA context is dead:", e, flush=True)
                return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
