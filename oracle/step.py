"""CPU fp32 restatement of SSRESRGANModel.feed_data / optimize_parameters
(/root/reference/ssr/models/ssr_esrgan_model.py:104-233) over the functional nets of oracle/nets.py.

State is explicit: parameter dicts for G, D, EMA, torch.optim.Adam instances over leaf tensors.  The statement order,
the three discriminator forwards (each advancing the spectral-norm power iteration), the loss weights and the
detach points follow the reference line by line; line numbers are quoted at each step.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import losses, nets


class OracleESRGAN:
    def __init__(self, g_params, d_params, vgg_params, opt, num_block=23):
        self.opt = opt
        self.num_block = num_block
        self.g = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in g_params.items())
        self.d = OrderedDict()
        for k, v in d_params.items():
            leaf = not (k.endswith("weight_u") or k.endswith("weight_v"))
            self.d[k] = v.clone().requires_grad_(True) if leaf else v.clone()
        self.vgg = vgg_params
        self.ema_decay = opt.get("ema_decay", 0.999)
        self.g_ema = OrderedDict((k, v.detach().clone()) for k, v in self.g.items())       # model_ema(0), :49
        kw = dict(lr=opt.get("lr", 1e-4), betas=tuple(opt.get("betas", (0.9, 0.99))), weight_decay=opt.get("weight_decay", 0))
        kw_d = dict(lr=opt.get("lr_d", kw["lr"]), betas=tuple(opt.get("betas_d", kw["betas"])),
                    weight_decay=opt.get("weight_decay_d", kw["weight_decay"]))            # train.optim_d, yml:103-107
        self.optimizer_g = torch.optim.Adam(list(self.g.values()), **kw)                      # :101 setup_optimizers
        self.optimizer_d = torch.optim.Adam([v for v in self.d.values() if v.requires_grad], **kw_d)
        self.net_d_iters = opt.get("net_d_iters", 1)                                         # :97-98
        self.net_d_init_iters = opt.get("net_d_init_iters", 0)
        self.scale = opt.get("scale", 4)
        self.feed_disc_lr = opt.get("feed_disc_lr", True)
        self.pixel_weight = opt.get("pixel_weight", 1.0)
        self.gan_weight = opt.get("gan_weight", 0.1)
        self.ssim_weight = opt.get("ssim_weight", 0.0)                                       # train.ssim_opt, :87-90
        self.percep = opt.get("perceptual", True)
        self.layer_weights = opt.get("layer_weights", losses.DEFAULT_LAYER_WEIGHTS)

    def feed_data(self, lr_u8, hr_u8, old_hr_u8=None):
        """:104-117"""
        with torch.no_grad():
            self.lr = lr_u8.float() / 255
            self.gt = hr_u8.float() / 255
            self.gt_usm = losses.usm_sharp(self.gt)
            self.old_hr = old_hr_u8.float() / 255 if old_hr_u8 is not None else None           # :112-114

    def net_g(self, x, params=None):
        return nets.rrdbnet_forward(params or self.g, x, scale=self.scale, num_block=self.num_block)

    def _disc_input(self, img, lr_resized):
        """:171-178 / :202-213: [img | lr_resized | old_hr]"""
        parts = [img]
        if self.feed_disc_lr:
            parts.append(lr_resized)
        if self.old_hr is not None:
            parts.append(self.old_hr)
        return torch.cat(parts, 1) if len(parts) > 1 else img

    def net_d(self, x):
        return nets.unet_disc_forward(self.d, x, training=True)

    def optimize_parameters(self, current_iter=1):
        """:119-233 with l1_gt_usm = percep_gt_usm = True, gan_gt_usm = False (esrgan_s2naip_urban.yml:9-11)."""
        o = self.opt
        do_g = (current_iter % self.net_d_iters == 0) and (current_iter > self.net_d_init_iters)   # :142
        l1_gt = self.gt_usm if o.get("l1_gt_usm", True) else self.gt
        percep_gt = self.gt_usm if o.get("percep_gt_usm", True) else self.gt
        gan_gt = self.gt_usm if o.get("gan_gt_usm", False) else self.gt
        lr_resized = F.interpolate(self.lr, scale_factor=4) if self.feed_disc_lr else None   # :133
        d_leaves = [v for v in self.d.values() if v.is_leaf and v.dtype.is_floating_point and v.grad_fn is None]
        for k, v in self.d.items():                                                          # :136-137
            if not (k.endswith("weight_u") or k.endswith("weight_v")):
                v.requires_grad_(False)
        self.optimizer_g.zero_grad()                                                         # :139
        self.output = self.net_g(self.lr)                                                    # :140
        log = OrderedDict()
        if do_g:
            l_g_total = 0
            l_g_pix = losses.l1_loss(self.output, l1_gt, self.pixel_weight)                  # :148
            l_g_total = l_g_total + l_g_pix
            log["l_g_pix"] = l_g_pix
            if self.percep:
                l_g_percep = losses.perceptual_loss(self.vgg, self.output, percep_gt, self.layer_weights)   # :154
                l_g_total = l_g_total + l_g_percep
                log["l_g_percep"] = l_g_percep
            if self.ssim_weight:
                l_g_ssim = losses.ssim_loss(self.output, percep_gt, self.ssim_weight)        # :163-166
                l_g_total = l_g_total + l_g_ssim
                log["l_g_ssim"] = l_g_ssim
            disc_input = self._disc_input(self.output, lr_resized)                           # :171-178
            fake_g_pred = self.net_d(disc_input)                                             # :181
            l_g_gan = losses.gan_loss_vanilla(fake_g_pred, True, is_disc=False, loss_weight=self.gan_weight)   # :182
            l_g_total = l_g_total + l_g_gan
            log["l_g_gan"] = l_g_gan
            l_g_total.backward()                                                             # :192
            self.optimizer_g.step()                                                          # :193
        for k, v in self.d.items():                                                          # :196-197
            if not (k.endswith("weight_u") or k.endswith("weight_v")):
                v.requires_grad_(True)
        fake_in = self._disc_input(self.output, lr_resized)                                  # :202-213
        real_in = self._disc_input(gan_gt, lr_resized)
        self.optimizer_d.zero_grad()                                                         # :215
        real_d_pred = self.net_d(real_in)                                                    # :217
        l_d_real = losses.gan_loss_vanilla(real_d_pred, True, is_disc=True)                  # :218
        log["l_d_real"] = l_d_real
        log["out_d_real"] = real_d_pred.detach().mean()
        l_d_real.backward()                                                                  # :221
        fake_d_pred = self.net_d(fake_in.detach().clone())                                   # :223
        l_d_fake = losses.gan_loss_vanilla(fake_d_pred, False, is_disc=True)                 # :224
        log["l_d_fake"] = l_d_fake
        log["out_d_fake"] = fake_d_pred.detach().mean()
        l_d_fake.backward()                                                                  # :227
        self.optimizer_d.step()                                                              # :228
        with torch.no_grad():                                                                # :230-231 model_ema
            for k in self.g:
                self.g_ema[k].mul_(self.ema_decay).add_(self.g[k].detach(), alpha=1 - self.ema_decay)
        self.log_dict = OrderedDict((k, float(v.detach())) for k, v in log.items())          # :233
        return self.log_dict
