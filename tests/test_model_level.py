"""Model level: the UNMODIFIED reference model (lit_gpt.model.GPT from baseline/_ref, see oracle/ref_loader.py)
with its attention dispatch rebound to the CUDA path by ``patch_causal_self_attention`` -- the drop-in of
SURVEY.md 8(b3).  The reference's own reshape after the dispatch (model.py:453-455, quirk Q1: no head
transpose) stays in force because ``CausalSelfAttention.forward`` is untouched.

CPU: the patch applies to the real class and the meta-device forward of the pretrain scripts' flop count
(pretrain/openwebtext.py:146-156, quirk Q7) runs through it.  GPU: logits / loss / gradients of Pythia-14m at
seq 256 against the fixture minted from the reference on the host (oracle/make_golden.py --only=model)."""
import pytest
import torch

from conftest import golden_model_names, load_golden_model
from oracle import ref_loader as rl

needs_ref = pytest.mark.skipif(not (rl.installed() or __import__("os").path.isdir(rl.REF_SRC)),
                               reason="reference not installed under baseline/_ref (run __graft_entry__.build() in the build container)")


@needs_ref
def test_patch_rebinds_the_real_dispatch_class_and_meta_forward_runs():
    from fastmax_experiments_b200.attention import patch_causal_self_attention
    lit_gpt = rl.reference_model()
    from lit_gpt.model import CausalSelfAttention
    orig = {n: getattr(CausalSelfAttention, n) for n in ("fastmax", "fastmax_cuda", "linearmax")}
    try:
        patch_causal_self_attention(CausalSelfAttention)
        assert CausalSelfAttention.fastmax is not orig["fastmax"]
        cfg = lit_gpt.Config.from_name("pythia-14m", block_size=128, attn_alg="fastmax")
        with torch.device("meta"):
            model = lit_gpt.GPT(cfg)
            logits = model(torch.zeros(2, 128, dtype=torch.long))
        assert logits.shape == (2, 128, cfg.padded_vocab_size) and logits.device.type == "meta"
        # the LoRA variant inherits the rebound methods (finetune/lora.py builds lit_gpt.lora.GPT)
        from lit_gpt import lora as lit_lora
        assert lit_lora.CausalSelfAttention.fastmax is CausalSelfAttention.fastmax
    finally:
        for n, f in orig.items():
            setattr(CausalSelfAttention, n, f)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", golden_model_names())
def test_reference_model_with_cuda_dispatch_matches_host_reference(name):
    from fastmax_experiments_b200.attention import patch_causal_self_attention
    gd = load_golden_model(name)
    rl.reference_model()
    from lit_gpt.model import CausalSelfAttention
    orig = {n: getattr(CausalSelfAttention, n) for n in ("fastmax", "fastmax_cuda", "linearmax")}
    try:
        model, cfg = rl.build_gpt(gd["model"], attn_alg="fastmax", block_size=gd["T"], seed=gd["seed"])
        patch_causal_self_attention(CausalSelfAttention)
        model = model.cuda()
        idx = gd["idx"].cuda()
        logits = model(idx)
        loss = torch.nn.functional.cross_entropy(logits.view(-1, logits.size(-1)), idx.view(-1))
        loss.backward()
        torch.cuda.synchronize()
        assert abs(float(loss) - gd["loss"]) <= 1e-4 * abs(gd["loss"])
        lse = torch.logsumexp(logits.detach().double(), -1).cpu()
        assert (lse - gd["lse"]).abs().max() <= 1e-3
        sample = logits.detach()[:, ::8, ::64].float().cpu()
        ref = gd["logits_sample"]
        assert (sample - ref).abs().max() <= 1e-3 * ref.abs().max() + 1e-4
        gnorm = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters() if p.grad is not None))
        assert abs(float(gnorm) - gd["grad_norm"]) <= 1e-3 * gd["grad_norm"]
        wte = model.transformer.wte.weight.grad[idx[0, :8]].float().cpu()
        assert (wte - gd["wte_grad_sample"]).abs().max() <= 1e-3 * gd["wte_grad_sample"].abs().max() + 1e-6
    finally:
        for n, f in orig.items():
            setattr(CausalSelfAttention, n, f)
