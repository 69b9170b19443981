"""Inference-side model wrapper with the reference's interface
(config/daclip-sde/models/{__init__,denoising_model,base_model,networks}.py of the reference):
`create_model(opt)` -> object with `.device`, `.model`, `.feed_data`, `.test`, `.output`,
`.get_current_visuals`, `.load`.  Accepts the reference's option dict / YAML keys unchanged
(`network_G.{which_model_G,setting}`, `path.{pretrain_model_G,strict_load}`, `gpu_ids`).
Training members (optimizers, EMA, schedulers, loss) are out of scope.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
from torch.nn.parallel import DataParallel, DistributedDataParallel

from . import unet as _unet

MODULES = {"ConditionalUNet": _unet.ConditionalUNet}


def define_G(opt):
    """networks.py:10-15: build `which_model_G(**setting)`."""
    net = opt["network_G"]
    return MODULES[net["which_model_G"]](**net["setting"])


class DenoisingModel:
    def __init__(self, opt):
        self.opt = opt
        self.device = torch.device("cuda" if opt.get("gpu_ids") is not None else "cpu")   # base_model.py:12
        self.is_train = bool(opt.get("is_train"))
        if self.is_train:
            raise NotImplementedError("daclip_b200 implements the inference path only")
        self.rank = -1
        self.model = define_G(opt).to(self.device)
        # single-process wrapper as in denoising_model.py:41-42; pinned to ONE device: multi-GPU runs shard the
        # batch one process per GPU instead of replicating the module every denoiser call
        self.model = DataParallel(self.model, device_ids=[torch.cuda.current_device()])
        self.load()
        self.state = self.condition = self.state_0 = self.output = None
        self.text_context = self.image_context = None

    # ---------------------------------------------------------------- checkpoints (base_model.py:92-105)
    def load_network(self, load_path, network, strict=True):
        if isinstance(network, (nn.DataParallel, DistributedDataParallel)):
            network = network.module
        self.load_state_dict_into_model(torch.load(load_path, map_location="cpu"), strict, network)

    def load_state_dict_into_model(self, state_dict, strict=True, network=None):
        network = network if network is not None else self.model.module
        clean = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in state_dict.items())
        network.load_state_dict(clean, strict=strict)

    def load(self):
        path = (self.opt.get("path") or {}).get("pretrain_model_G")
        if path is not None:
            self.load_network(path, self.model, (self.opt.get("path") or {}).get("strict_load", True))

    # ---------------------------------------------------------------- inference API (denoising_model.py:121-173)
    def feed_data(self, state, LQ, GT=None, text_context=None, image_context=None):
        self.state = state.to(self.device)
        self.condition = LQ.to(self.device)
        if GT is not None:
            self.state_0 = GT.to(self.device)
        self.text_context = text_context
        self.image_context = image_context

    def test(self, sde=None, mode="posterior", save_states=False):
        sde.set_mu(self.condition)
        self.model.eval()
        with torch.no_grad():
            if mode == "sde":
                self.output = sde.reverse_sde(self.state, save_states=save_states, text_context=self.text_context,
                                              image_context=self.image_context)
            else:
                self.output = sde.reverse_posterior(self.state, save_states=save_states,
                                                    text_context=self.text_context, image_context=self.image_context)
        self.model.train()

    def get_current_visuals(self, need_GT=True):
        out = OrderedDict()
        out["Input"] = self.condition.detach()[0].float().cpu()
        out["Output"] = self.output.detach()[0].float().cpu()
        if need_GT:
            out["GT"] = self.state_0.detach()[0].float().cpu()
        return out


def create_model(opt):
    """models/__init__.py:6-15."""
    if opt.get("model", "denoising") != "denoising":
        raise NotImplementedError(f"Model [{opt.get('model')}] not recognized.")
    return DenoisingModel(opt)
