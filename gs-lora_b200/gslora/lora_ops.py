"""Stand-alone forward of a loralib layer (outside ViT_face's fused engine) through the C-ABI GEMM family.
Inside `ViT_face` these layers are never called: the engine runs the whole block."""
import torch

from . import _ffi as F


def lora_linear_forward(layer, x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise F.GslError("gslora-b200: loralib layers execute on CUDA (sm_100a) only; there is no CPU fallback")
    if torch.is_grad_enabled() and any(p.requires_grad for p in layer.parameters()):
        raise NotImplementedError("gslora-b200: stand-alone training through a single loralib layer is not built yet; "
                                  "use vit_pytorch_face.ViT_face (fused engine)")
    in_f, out_f = layer.in_features, layer.out_features
    r = getattr(layer, "r", 0)
    use_lora = r > 0 and not layer.merged
    lead = x.shape[:-1]
    x2 = x.reshape(-1, in_f)
    M = x2.shape[0]
    xcat = torch.zeros(M, in_f + 16, dtype=torch.half, device=x.device)
    xcat[:, :in_f] = x2
    wcat = torch.zeros(out_f, in_f + 16, dtype=torch.half, device=x.device)
    wcat[:, :in_f] = layer.weight.data
    K = in_f
    if use_lora:
        a16 = torch.zeros(16, in_f, dtype=torch.half, device=x.device)
        a16[:r] = layer.lora_A.data
        F.check(F.lib().gsl_lora_down(F.ptr(xcat), in_f + 16, F.ptr(a16), in_f, F.ptr(xcat[:, in_f:]), in_f + 16, M, in_f, 8 if r <= 8 else 16,
                                      F.cur_stream()), "gsl_lora_down")
        wcat[:, in_f:in_f + r] = layer.lora_B.data * layer.scaling
        K = in_f + 16
    out = torch.empty(M, out_f, dtype=torch.float32, device=x.device)
    F.gemm_f16(xcat, wcat, epi=F.EPI_F32, bias=layer.bias.data if layer.bias is not None else None, out0=out, K=K)
    return out.reshape(*lead, out_f)
