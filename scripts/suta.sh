#!/usr/bin/env bash
# The experiment presets of the reference (REF/scripts/LS.sh, CH.sh, CV.sh, TD.sh: 10-step EM+MCC SUTA with --train_feature,
# episodic, lr 2e-5, temperature 2.5, non-blank frames, reweighting) as ONE parametrised launcher over this repo's main.py.
#
#   scripts/suta.sh <corpus> <dataset_dir> [noise ...] [-- extra main.py flags]
#
#   corpus       LS | CH | CV | TD        -> --dataset_name librispeech | chime | commonvoice | ted
#   dataset_dir  LibriSpeech root, CHiME3 root, cv-corpus-*/en, TEDLIUM_release2/test (the layouts of REF/corpus/*.py)
#   noise        levels of REF/data.py:23's Gaussian noise, one run each (default: LS "0 0.005 0.01" like LS.sh, others "0")
#
# Environment: ASR (default facebook/wav2vec2-base-960h: a hub name needs the network once; a local HF directory or
# random-base works offline), BATCH_UTTS (default 64: utterances adapted per step on the GPU; 0 = the reference's
# one-utterance-at-a-time loop), LOG_DIR (default exps), DRY_RUN=1 (print the commands, run nothing).
set -euo pipefail
if [ $# -lt 2 ]; then sed -n 2,13p "$0"; exit 2; fi
corpus=$1; dir=$2; shift 2
case "$corpus" in
  LS) name=librispeech; levels="0 0.005 0.01" ;;
  CH) name=chime; levels="0" ;;
  CV) name=commonvoice; levels="0" ;;
  TD) name=ted; levels="0" ;;
  *) echo "unknown corpus '$corpus' (LS, CH, CV or TD)" >&2; exit 2 ;;
esac
picked=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do picked+=("$1"); shift; done
[ $# -gt 0 ] && shift
[ ${#picked[@]} -gt 0 ] && levels="${picked[*]}"
here=$(cd "$(dirname "$0")/.." && pwd)
batch=${BATCH_UTTS:-64}
for noise in $levels; do
  cmd=(python "$here/main.py" --asr "${ASR:-facebook/wav2vec2-base-960h}" --dataset_name "$name" --dataset_dir "$dir"
       --steps 10 --episodic --lr 2e-5 --temp 2.5 --em_coef 0.3 --reweight --non_blank --train_feature
       --extra_noise "$noise" --log_dir "${LOG_DIR:-exps}")
  [ "$batch" -gt 1 ] && cmd+=(--batch_utts "$batch")
  echo "+ ${cmd[*]} $*" >&2
  [ -n "${DRY_RUN:-}" ] || "${cmd[@]}" "$@"
done
