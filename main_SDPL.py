"""Drop-in replacement for the reference's main_SDPL.py (the README's SDPL baseline: adaptation by a CTC pseudo-label
loss, REF/main_SDPL.py): same function names and signatures, same CLI flags (--pl_coef, Adam, 10 steps by default), same
result file -- on the B200 engine.  The pseudo-label loss itself is csrc/ctc.cu."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "test-time-adaptation-asr-suta_b200"))

from suta_b200.api import (configure_model, copy_model_and_optimizer, div_loss, load_model_and_optimizer,  # noqa: E402,F401
                           mcc_loss, softmax_entropy, SutaModel)
from suta_b200.api import sdpl_collect_params as collect_params  # noqa: E402,F401
from suta_b200.api import sdpl_forward_and_adapt as forward_and_adapt  # noqa: E402,F401
from suta_b200.api import sdpl_setup_optimizer as setup_optimizer  # noqa: E402,F401
from suta_b200.api import pseudo_labeling_loss  # noqa: E402,F401
from suta_b200.wer import wer  # noqa: E402,F401

if __name__ == '__main__':
    from suta_b200.cli import main
    main(sdpl=True)
