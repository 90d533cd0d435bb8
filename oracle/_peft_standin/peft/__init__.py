"""Minimal stand-in for the `peft` package (not installed offline) so that the
reference's modelling_longitudinal.py imports and applies LoRA the way
`get_peft_model` would: every module whose qualified name FULL-matches
`target_modules` (peft uses re.fullmatch for string targets) is replaced by a
LoRA-wrapped Linear computing  base(x) + lora_B(lora_A(dropout(x))) * alpha / r.

TEST INFRASTRUCTURE: used only by oracle/pin_against_reference.py.
"""
import re
from dataclasses import dataclass
from enum import Enum

import torch


class TaskType(str, Enum):
    CAUSAL_LM = "CAUSAL_LM"


@dataclass
class LoraConfig:
    inference_mode: bool = False
    r: int = 8
    lora_alpha: int = 8
    lora_dropout: float = 0.0
    target_modules: object = None


def get_peft_config(d):
    return LoraConfig(**d)


class _Adapters(torch.nn.ModuleDict):
    pass


class LoraLinear(torch.nn.Module):
    def __init__(self, base: torch.nn.Linear, cfg: LoraConfig):
        super().__init__()
        self.base_layer = base
        self.lora_A = _Adapters({"default": torch.nn.Linear(base.in_features, cfg.r, bias=False)})
        self.lora_B = _Adapters({"default": torch.nn.Linear(cfg.r, base.out_features, bias=False)})
        self.lora_dropout = torch.nn.Dropout(cfg.lora_dropout)
        self.scaling = cfg.lora_alpha / cfg.r
        torch.nn.init.zeros_(self.lora_B["default"].weight)
        for p in base.parameters():
            p.requires_grad = False

    def forward(self, x):
        return self.base_layer(x) + self.lora_B["default"](self.lora_A["default"](self.lora_dropout(x))) * self.scaling


class _LoraModel(torch.nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model


class PeftModel(torch.nn.Module):
    def __init__(self, model, cfg: LoraConfig):
        super().__init__()
        for p in model.parameters():
            p.requires_grad = False
        targets = [n for n, m in model.named_modules()
                   if isinstance(m, torch.nn.Linear) and re.fullmatch(cfg.target_modules, n)]
        for name in targets:
            parent_name, _, child = name.rpartition(".")
            parent = model.get_submodule(parent_name)
            setattr(parent, child, LoraLinear(getattr(parent, child), cfg))
        self.base_model = _LoraModel(model)
        self.n_targets = len(targets)

    def forward(self, *a, **k):
        return self.base_model.model(*a, **k)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.base_model.model, name)

    def print_trainable_parameters(self):
        tr = sum(p.numel() for p in self.parameters() if p.requires_grad)
        al = sum(p.numel() for p in self.parameters())
        print(f"trainable params: {tr} || all params: {al} || trainable%: {100 * tr / al}")


def get_peft_model(model, cfg):
    return PeftModel(model, cfg)
