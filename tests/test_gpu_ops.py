"""GPU: every CUDA operator, called through the C ABI, against fp32 torch / the oracle (tolerances stated here).
fp32 outputs of bf16-operand GEMMs must match an fp32 reference on the same bf16 inputs to accumulation-order noise
(1e-5 relative); bf16 outputs carry one bf16 rounding (2^-9 relative per element, < 4e-3 in norm)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

F32, BF16 = 1e-5, 4e-3


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import gpu_checks
    return gpu_checks


def test_gemm_plain(K):
    r = K.check_gemm_plain()
    assert r["nan"] == 0 and r["rel"] < F32


# N = 384 runs the one-SM kernel; N = 768 (N % 256 == 0) the CTA-pair kernel: 5 / 11 row blocks = an odd number of pairs' halves,
# ragged last block, K not a multiple of the ring depth
@pytest.mark.parametrize("M,N,Kd", [(517, 384, 256), (517, 768, 256), (1300, 512, 832), (129, 256, 64)])
def test_gemm_epilogues(K, M, N, Kd):
    r = K.check_gemm_epilogue(M, N, Kd, seed=M + N)
    assert r["res_rel"] < F32 and r["gelu_bwd_rel"] < F32 and r["acc_rel"] < F32
    assert r["bf16_rel"] < BF16 and r["gelu_bwd16_rel"] < BF16
    assert r["gelu_rel"] < BF16 and r["aux_rel"] < BF16 and r["res16_rel"] < BF16 and r["gelu_bwd_vs_exact_rel"] < BF16


def test_gemm_pair_kernel_tail_split(K):
    r = K.check_gemm_pair_tail()                     # 31 row pairs x 3 column blocks = 93 tiles = 74 + 19 -> 3-way K split
    assert r["acc_rel"] < F32 and r["bf16_rel"] < BF16
    r = K.check_gemm_pair_tail(M=20000, N=768, K=3072, seed=22)
    assert r["acc_rel"] < F32 and r["bf16_rel"] < BF16


def test_gemm_shapes_of_the_path(K):
    for name, rel in K.check_gemm_shapes().items():
        assert rel < 2e-5, name


def test_gemm_overlapping_window_operand(K):
    assert K.check_gemm_window()["rel"] < 2e-5
    assert K.check_gemm_window(R=300, CG=64, Kp=16, N=64)["rel"] < 2e-5
    assert K.check_gemm_strided_conv()["rel"] < F32
    assert K.check_gemm_strided_conv(L=4003, C=512, k=2, s=2, N=512)["rel"] < F32


@pytest.mark.parametrize("N,bf16_in", [(768, False), (512, True), (1024, False), (128, False), (64, True)])
def test_layernorm_forward_backward(K, N, bf16_in):
    r = K.check_layernorm(N, (70, 3, 129, 64), 5 + N, bf16_in)
    assert r["y_rel"] < F32 and r["dx_rel"] < F32 and r["dparam_rel"] < F32
    assert r["y16_rel"] < BF16 and r["dx16_rel"] < BF16
    # dgamma/dbeta: fixed-order two-stage reduction (no atomics) -> the same bits on every run, nothing else touched
    assert r["dparam_bit_equal"] and r["dparam_writes_only_its_segment"]


@pytest.mark.parametrize("N,dy_bf16", [(512, True), (512, False), (64, True)])
def test_layernorm_gelu_conv_layout(K, N, dy_bf16):
    """lv60 conv layers: GELU(LayerNorm(x)), gap rows skipped, bf16 gradient replaced in place (HF:291-299)."""
    r = K.check_layernorm_gelu(N, (70, 3, 129, 300), (5, 0, 250, 13), 9 + N, dy_bf16)
    assert r["nan"] == 0 and r["gaps_untouched"] and r["dx_gaps_zero"], r
    # GELU / GELU' use the engine's erf polynomial (common.cuh: ~1e-6 absolute): 1e-4 instead of the 1e-5 of plain LayerNorm
    assert r["y16_rel"] < BF16 and r["dparam_rel"] < 1e-4, r
    assert r["dx_rel"] < (BF16 if dy_bf16 else 1e-4), r


def test_layernorm_gelu_many_rows(K):
    r = K.check_layernorm_gelu(512, (5000, 1, 3, 12000, 700), (250, 255, 1, 0, 100), 78, True)
    assert r["nan"] == 0 and r["dx_rel"] < BF16 and r["dparam_rel"] < 1e-4 and r["gaps_untouched"], r


@pytest.mark.parametrize("N", [768, 1024, 128])
def test_layernorm_keep_input_and_residual_backward(K, N):
    """lv60 pre-LN encoder: the LayerNorm seeds the residual buffer with its input + bias; its backward adds onto d residual."""
    r = K.check_layernorm_keep_input(N, (70, 3, 129, 64), 10 + N)
    assert r["y32_rel"] < F32 and r["dx_rel"] < F32 and r["dparam_rel"] < F32, r
    assert r["y16_rel"] < BF16 and r["dx16_rel"] < BF16, r


def test_layernorm_backward_many_rows_many_utterances(K):
    """Several CTAs per utterance and several utterances per CTA (the slot = CTA + utterance bookkeeping)."""
    r = K.check_layernorm(768, (5000, 1, 1, 3, 12000, 2, 700), 77, False)
    assert r["dx_rel"] < F32 and r["dparam_rel"] < F32 and r["dparam_bit_equal"]


@pytest.mark.parametrize("Ts,heads", [((249, 64, 1, 130), 2), ((1749,), 1), ((63, 65, 128, 129), 12)])
def test_attention_forward_backward(K, Ts, heads):
    r = K.check_attention(Ts, heads)
    assert r["nan"] == 0
    assert r["o_rel"] < 6e-3 and r["dq_rel"] < 8e-3 and r["dk_rel"] < 8e-3 and r["dv_rel"] < 8e-3


@pytest.mark.parametrize("em,rew,nb", [(0.3, True, True), (0.3, False, False), (1.0, False, True), (0.0, True, False)])
def test_fused_loss_matches_reference_formulas(K, em, rew, nb):
    r = K.check_loss(em_coef=em, reweight=rew, not_blank=nb)
    assert r["loss_rel"] < 1e-5 and r["grad_rel"] < 1e-4 and r["bf16_rel"] < BF16


def test_fused_loss_single_frame_and_longest_utterance(K):
    r = K.check_loss(Ts=(1, 2, 1874), blank_bias=0.5)       # a one-frame utterance whose frame is blank has a NaN loss in the reference too
    assert r["nan_where_the_reference_is_nan"] and r["loss_rel"] < 1e-5 and r["grad_rel"] < 1e-4


def test_fused_loss_all_frames_blank(K):
    """REF/main.py:190: with --non_blank and every frame predicted blank the entropy term is the mean of an empty
    selection = NaN.  The kernel must report that NaN loss (not hide it) and keep d loss / d logits finite."""
    r = K.check_loss(Ts=(40, 5, 1), blank_bias=50.0)
    assert r["nan_loss_utts"] == 3 and r["nan_where_the_reference_is_nan"] and r["grad_finite_where_loss_nan"]
    # strongly blank-dominated logits without the mask are an ordinary case
    r = K.check_loss(Ts=(40, 5, 1), blank_bias=6.0, not_blank=False)
    assert r["nan_loss_utts"] == 0 and r["loss_rel"] < 1e-5 and r["grad_rel"] < 1e-3


@pytest.mark.parametrize("em,rew,nb,div", [(0.3, True, True, 0.25), (1.0, False, False, 0.3), (0.0, True, True, 0.1)])
def test_fused_loss_with_div_loss(K, em, rew, nb, div):
    """--div_coef > 0 (REF/main.py:46-60,201-203): minus the entropy of the time-averaged non-blank logits."""
    r = K.check_loss(em_coef=em, reweight=rew, not_blank=nb, div_coef=div)
    assert r["loss_rel"] < 1e-5 and r["grad_rel"] < 1e-4


@pytest.mark.parametrize("Ts,seed,blank_utt", [((249, 37, 1, 6, 700), 10, 1), ((1874, 2), 11, None), ((64, 65, 33), 12, 2)])
def test_pseudo_label_ctc_loss_matches_torch_ctcloss(K, Ts, seed, blank_utt):
    """SDPL baseline (REF/main_SDPL.py:194-209): value and gradient of nn.CTCLoss against the greedy transcript, with the
    reference's log_softmax over TIME; empty targets (all frames blank), one-frame utterances, repeated characters, the
    longest utterance the data loader lets through (T = 1874)."""
    r = K.check_ctc(Ts, seed, all_blank_utt=blank_utt)
    assert r["target_len_equal"] and r["finite"] and r["bit_equal"]
    assert r["loss_rel"] < 1e-4 and r["grad_rel"] < 1e-3, r


def test_gemm_mn_major_operands(K):
    """Weight-gradient / per-utterance dgrad GEMMs read operands stored [K rows][M|N contiguous]."""
    for args in ((), (128, 64, 64, 14), (768, 512, 4000, 15), (512, 1536, 2000, 16)):
        for name, rel in K.check_gemm_mn(*args).items():
            assert rel < 2e-5, (args, name, rel)


def test_adam_with_multiplicities(K):
    r = K.check_adam()
    assert r["frozen_moved"] == 0.0 and r["delta_rel"] < 1e-4 and r["m_rel"] < 1e-5 and r["v_rel"] < 1e-4
    # --train_all lists an encoder Linear 7 times, --train_all --train_feature a conv weight 10 times (REF/main.py:88-100)
    r = K.check_adam(max_mult=12, seed=18)
    assert r["frozen_moved"] == 0.0 and r["delta_rel"] < 1e-4 and r["m_rel"] < 1e-5 and r["v_rel"] < 1e-4


def test_ctc_decode_bit_exact(K):
    r = K.check_decode()
    assert r == {"argmax_mismatch": 0, "collapse_mismatch": 0}
    assert K.check_decode(Ts=(1, 1, 1874, 2), seed=3) == {"argmax_mismatch": 0, "collapse_mismatch": 0}
