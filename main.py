"""Drop-in replacement for the reference's main.py (REF/main.py): same function names, same CLI flags, same
per-utterance loop and result files -- the arithmetic runs on the B200 engine (suta_b200).  See suta_b200/cli.py for the
offline differences (`--asr random-base|random-large|<local dir>`, `--dataset_name synthetic`) and the `--batch_utts N`
extension."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "test-time-adaptation-asr-suta_b200"))

from suta_b200.api import (collect_params, configure_model, copy_model_and_optimizer, div_loss,  # noqa: E402,F401
                           forward_and_adapt, load_model_and_optimizer, mcc_loss, setup_optimizer, softmax_entropy,
                           SutaModel)
from suta_b200.wer import wer  # noqa: E402,F401

if __name__ == '__main__':
    from suta_b200.cli import main
    main(sdpl=False)
