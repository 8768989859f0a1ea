"""A tiny decoder-like module tree shared by the golden-checkpoint generator (which runs the REFERENCE's save_model on
it) and the tests (which load that checkpoint with this repo's loader).  No reference code is imported here."""
import torch

HIDDEN, FFN, NOUT, GROUP = 256, 512, 128, 128
QUANT_NAMES = ("self_attn.q_proj", "self_attn.o_proj", "mlp.down_proj")


class Attn(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.q_proj = torch.nn.Linear(HIDDEN, HIDDEN, bias=False)
        self.o_proj = torch.nn.Linear(HIDDEN, HIDDEN, bias=False)


class Mlp(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.down_proj = torch.nn.Linear(FFN, HIDDEN, bias=True)


class Block(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.self_attn = Attn()
        self.mlp = Mlp()
        self.input_layernorm = torch.nn.LayerNorm(HIDDEN)


class TinyModel(torch.nn.Module):
    def __init__(self, nblocks=2):
        super().__init__()
        self.layers = torch.nn.ModuleList([Block() for _ in range(nblocks)])
        self.lm_head = torch.nn.Linear(HIDDEN, 64, bias=False)
        self.dtype = torch.float16


def build(seed=0):
    torch.manual_seed(seed)
    m = TinyModel().half()
    m.dtype = torch.float16
    return m


def quant_layer_names(nblocks=2):
    return [f"layers.{i}.{n}" for i in range(nblocks) for n in QUANT_NAMES]
