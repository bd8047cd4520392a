"""-m gpu: every C-ABI entry point against fp32 PyTorch / the CPU oracle on seeded inputs."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_checks
    return gpu_checks


def test_pack_roundtrip(G):
    G.check_pack_roundtrip()


@pytest.mark.parametrize("ksize,nsrc,B,H,W", [(3, 1, 1, 32, 24), (3, 3, 2, 37, 45), (3, 5, 1, 48, 40), (1, 3, 2, 37, 45),
                                               (3, 2, 1, 5, 3), (3, 2, 2, 144, 180)])
def test_conv_tcgen05_tap(G, ksize, nsrc, B, H, W):
    G.check_conv(impl=G.K.IMPL_TAP, ksize=ksize, nsrc=nsrc, B=B, H=H, W=W)


@pytest.mark.parametrize("nsrc,ngroups,B", [(1, 6, 3), (2, 3, 3), (3, 2, 4), (5, 1, 2)])
def test_conv_repeatability_at_bench_shape(G, nsrc, ngroups, B):
    G.check_conv_repeatability(nsrc=nsrc, ngroups=ngroups, B=B)


def test_conv_checker_kernel(G):
    G.check_conv(impl=G.K.IMPL_CHECK, ksize=3, nsrc=2, B=2, H=21, W=19)


def test_conv_aux16(G):
    G.check_conv_aux16()


def test_conv_per_sample_weights(G):
    G.check_osa_conv_per_sample()


def test_pack_frames_and_first_layer_on_tensor_cores(G):
    G.check_pack_frames()
    G.check_pack_frames(B=1, h=16, w=20, seed=16)


@pytest.mark.parametrize("ci,B", [(192, 2), (320, 1), (64, 3)])
def test_osa_prologue(G, ci, B):
    G.check_osa_prologue(ci=ci, B=B)


def test_channel_attention(G):
    G.check_ca()


def test_osadapt_mask(G):
    G.check_mask()


@pytest.mark.parametrize("h,w,scale", [(144, 180, (4, 4)), (144, 180, (1.5, 4)), (144, 180, (2.7, 2.7)), (64, 64, (2, 2)),
                                         (180, 318, (4, 4)), (63, 65, (2.7, 2.7)), (33, 35, (1.5, 1.5)), (31, 31, (3, 3))])
def test_satu_index_bit_exact(G, h, w, scale):
    G.check_satu_index(h, w, scale)


def test_satu_table(G):
    G.check_satu_table()


def test_satu_kernel_conv_and_sta_fused(G):
    G.check_satu_kconv_sta()
    G.check_satu_kconv_sta(B=3, h=32, w=40, seed=3)      # several tiles per sample, even tile count
    G.check_satu_kconv_sta(B=1, h=17, w=9, seed=4)       # odd sizes: padded column / row, single-tile batches
    G.check_satu_kconv_sta(B=3, h=16, w=24, seed=5)      # 3 tiles per sample: a single-tile batch between samples
    G.check_satu_kconv_sta(B=2, h=48, w=40, seed=6)      # 15 tiles per sample, several CTAs, odd count


@pytest.mark.parametrize("kw", [dict(), dict(B=1, h=16, w=20, scale=(4, 4), seed=2), dict(B=3, h=36, w=45, scale=(4, 4), seed=3, offset_gain=1.0),
                                dict(B=2, h=30, w=22, scale=(1.5, 4), seed=4), dict(B=1, h=9, w=70, scale=(1.1, 1.2), seed=5),
                                dict(B=2, h=21, w=33, scale=(3, 3), seed=6), dict(B=1, h=2, w=3, scale=(8, 7.3), seed=7)])
def test_satu_hr_one_kernel(G, kw):
    G.check_satu_hr(**kw)


@pytest.mark.parametrize("kw", [dict(B=2, nsrc=1, H=16, W=20), dict(B=2, nsrc=2, H=16, W=20), dict(B=1, nsrc=3, H=13, W=15, seed=1),
                                dict(B=3, nsrc=3, H=18, W=26, per_sample=True, seed=2), dict(B=2, nsrc=5, H=9, W=70, seed=3),
                                dict(B=4, nsrc=2, H=64, W=64, seed=4), dict(B=2, nsrc=1, H=33, W=130, per_sample=True, seed=5),
                                dict(B=1, nsrc=1, H=1, W=3, bias=False, seed=6)])
def test_conv3x3_autograd_dgrad_wgrad(G, kw):
    """Row f1 stage A: backward of the dominant op on tcgen05 (dgrad = forward kernel on flipped weights, wgrad = pixel contraction)."""
    G.check_conv_autograd(**kw)


def test_device_tensor2img_and_psnr_y(G):
    G.check_img_metrics()


def test_device_lr_synthesis_bit_exact(G):
    G.check_lr_synthesis()


def test_device_clip_evaluation_loop(G):
    G.check_evaluate_clip()


def test_clip_prefetcher_and_dataset_summary(G):
    G.check_clip_prefetch()


def test_kernels_in_fp16_operand_format(G):
    """The same entry points with the context switched to the fp16 storage / operand format."""
    G.set_precision("fp16")
    try:
        G.check_pack_roundtrip()
        G.check_conv(impl=G.K.IMPL_HALO, ksize=3, nsrc=2, B=2, H=37, W=45)
        G.check_conv(impl=G.K.IMPL_HALO, ksize=3, nsrc=3, B=1, H=48, W=40)
        G.check_conv(impl=G.K.IMPL_TAP, ksize=1, nsrc=3, B=2, H=37, W=45)
        G.check_osa_conv_per_sample(impl=G.K.IMPL_HALO)
        G.check_pack_frames()
        G.check_osa_prologue(ci=192, B=2)
        G.check_ca()
        G.check_satu_kconv_sta()
        G.check_satu_hr()
    finally:
        G.set_precision("bf16")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 5, 7, 9), (1, 3, 2, 3), (1, 2, 1, 6), (2, 64, 16, 20)])
def test_sta_lrelu_matches_the_aten_formulation(G, shape):
    """Training: kernel_conv's LeakyReLU + sta_conv fused (replicate padding, 25 taps), forward and gradients, incl. maps smaller than the 5x5 window."""
    B, C, h, w = shape
    print(G.check_sta_lrelu(B, C, h, w))


@pytest.mark.gpu
def test_training_weight_packing_table(G):
    """One table-driven launch packs every filter corner, forward orientation and transposed + flipped, bit-identical to the single-filter packer."""
    print(G.check_pack_chunks())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 6, 20), (1, 4, 72), (3, 5, 64)])
def test_training_elementwise_kernels(G, shape):
    """axpby / grad_prep / nchw3 against their definitions, with a padded pitch, a second 64-pixel chunk and the exact fit."""
    print(G.check_train_elementwise(*shape))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 10, 20), (1, 6, 72)])
def test_training_batched_weight_gradient(G, shape):
    print(G.check_wgrad_batched(*shape))


@pytest.mark.gpu
def test_training_adam_ema_kernel(G):
    print(G.check_adam_ema())
