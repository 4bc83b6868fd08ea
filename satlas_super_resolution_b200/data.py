"""Packed uint8 shards for the S2-NAIP training pairs -- the data format on the input side of `feed_data`.

The reference dataset (`ssr/data/s2-naip_dataset.py:34-249`) decodes one NAIP PNG and one PNG per Sentinel-2 band for every
sample, every epoch; at > 1 k img-pairs/s per GPU that PNG decode is the bottleneck of the whole step.  Here the PNGs are
decoded ONCE into a flat binary shard (`pack_s2naip`); `S2NAIPShardDataset` memory-maps the shard and makes the reference's
`__getitem__` decisions -- the black-pixel rejection of the NAIP chip (`:171-175`), the per-band assembly with zero tensors for
missing bands (`:180-198`), the clean / black frame split and `random.sample` choice (`:206-222`), the random-crop augmentation
(`:226-233`) and the `[T*C, 32, 32]` reshape (`:236-237`) -- from per-record flags computed at pack time, copying only the
chosen frames' bytes.  With the same `random` state it returns the SAME tensors and consumes the SAME random numbers as the
unmodified reference reading the PNG tree (`tests/test_data_cpu.py` compares item by item).  Without `shard_path` the class
decodes the PNG tree itself (the reference's behaviour; the baseline arm of `scripts/bench_ingest.py`: 1.1 k samples/s per
host thread vs 18 k from the shard).

Shard = `<prefix>.bin` (concatenated uint8 arrays) + `<prefix>.json` (per record: NAIP path, chip name, `hr_black`, and for
every stored array its offset and shape; the first band also carries the per-frame `black` flags).  The `lr` tensor keeps the
reference's channel order `t * C + c`, TCI RGB first.
"""
import glob
import json
import os
import random

import numpy as np
import torch
import torch.nn.functional as F
from torch.utils import data as data
from torch.utils.data import WeightedRandomSampler

from .registry import DATASET_REGISTRY, _register

SHARD_VERSION = 2


def has_black_pixels(tensor):
    """ssr/utils/data_utils.py:3-10 -- any pixel whose channel sum is 0"""
    return (torch.sum(tensor, dim=0).view(-1) == 0).any()


def _read_png(path):
    import torchvision
    return torchvision.io.read_image(path)


def _ordered_bands(opt):
    """`s2_bands` with 'tci' moved to the front (s2-naip_dataset.py:70-72); the option list is reordered in place like there"""
    bands = opt["s2_bands"] if "s2_bands" in opt else ["tci"]
    bands.insert(0, bands.pop(bands.index("tci")))
    return bands


def _old_naip_index(old_naip_path):
    chips = {}
    for old_naip in glob.glob(old_naip_path + "/**/*.png", recursive=True):
        chips.setdefault(old_naip.split("/")[-1][:-4], []).append(old_naip)
    return chips


def pack_s2naip(opt, out_prefix):
    """Decode the PNG tree named by `opt` (same keys as the reference dataset: sentinel2_path, naip_path, s2_bands,
    old_naip_path) once and write `<out_prefix>.bin` / `.json`.  Records follow the reference's glob order."""
    bands = _ordered_bands(opt)
    s2_root, naip_root = opt["sentinel2_path"], opt["naip_path"]
    old_index = _old_naip_index(opt["old_naip_path"]) if opt.get("old_naip_path") else None
    records, off = [], 0
    with open(out_prefix + ".bin", "wb") as fh:
        def put(t):
            nonlocal off
            a = np.ascontiguousarray(t.numpy())
            fh.write(a.tobytes())
            entry = {"off": off, "shape": list(a.shape)}
            off += a.size
            return entry

        for n in glob.glob(naip_root + "/**/*.png", recursive=True):
            chip = n.split("/")[-2]
            hr = _read_png(n)
            # the validity tests of the reader are functions of the pixels alone: evaluate them once, here
            rec = {"naip": n, "chip": chip, "hr": put(hr), "hr_black": bool(has_black_pixels(hr)), "bands": []}
            if old_index is not None:
                rec["old_hr"] = put(_read_png(old_index[chip][0]))
            for band in bands:
                p = os.path.join(s2_root, chip, band + ".png")
                if not os.path.exists(p):
                    rec["bands"].append({"missing": True, "tci": "tci" in p})
                    continue
                try:
                    img = _read_png(p)
                    # [C, T*32, 32] -> [T, C, 32, 32]  (s2-naip_dataset.py:189-190)
                    img = torch.reshape(img, (img.shape[0], -1, 32, 32)).permute(1, 0, 2, 3)
                    entry = put(img)
                    if not rec["bands"]:      # first band = TCI: which frames carry a black pixel (frame choice, :206-214)
                        entry["black"] = [bool(has_black_pixels(f[:3])) for f in img]
                    rec["bands"].append(entry)
                except Exception:
                    rec["bands"].append({"broken": True})
            records.append(rec)
    with open(out_prefix + ".json", "w") as fh:
        json.dump({"version": SHARD_VERSION, "bands": bands, "records": records}, fh)
    return len(records)


class CustomWeightedRandomSampler(WeightedRandomSampler):
    """s2-naip_dataset.py:18-31: weighted sampling that is not limited to 2^24 entries"""

    def __iter__(self):
        rand = np.random.choice(range(0, len(self.weights)), size=self.num_samples,
                                p=self.weights.numpy() / torch.sum(self.weights).numpy(), replace=self.replacement)
        return iter(torch.from_numpy(rand).tolist())


class S2NAIPShardDataset(data.Dataset):
    """Same `opt` keys and the same returned dict as the reference `S2NAIPDataset`; `shard_path` = prefix written by
    `pack_s2naip` (omit it to decode the PNG tree directly)."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.split = opt["phase"]
        train = self.split == "train"
        self.rand_crop = opt["rand_crop"] if "rand_crop" in opt else False
        self.n_s2_images = int(opt["n_s2_images"])
        self.scale = int(opt["scale"])
        self.use_3d = opt["use_3d"] if "use_3d" in opt else False
        self.old_naip_path = opt["old_naip_path"] if "old_naip_path" in opt else None
        if opt.get("osm_objs_path"):
            raise NotImplementedError("osm_objs_path: the OSM-object GAN variant is outside the built path (SURVEY.md section 2)")
        self.shard_path = opt.get("shard_path")
        if self.shard_path:
            with open(self.shard_path + ".json") as fh:
                idx = json.load(fh)
            if idx.get("version") != SHARD_VERSION:
                raise ValueError(f"{self.shard_path}.json: shard version {idx.get('version')} (expected {SHARD_VERSION})")
            want = _ordered_bands(opt)
            if idx["bands"] != want:
                raise ValueError(f"shard holds bands {idx['bands']}, the config asks for {want}")
            self.s2_bands = want
            if self.old_naip_path is not None and idx["records"] and "old_hr" not in idx["records"][0]:
                raise ValueError("old_naip_path is set but the shard was packed without it")
            self._bin = np.memmap(self.shard_path + ".bin", dtype=np.uint8, mode="r")
            records = idx["records"]
        else:
            self.s2_bands = _ordered_bands(opt)
            self.s2_path, self.naip_path = opt["sentinel2_path"], opt["naip_path"]
            if not (os.path.exists(self.s2_path) and os.path.exists(self.naip_path)):
                raise Exception("Please make sure the paths to the data directories are correct.")
            old_index = _old_naip_index(self.old_naip_path) if self.old_naip_path is not None else None
            records = []
            for n in glob.glob(self.naip_path + "/**/*.png", recursive=True):
                chip = n.split("/")[-2]
                rec = {"naip": n, "chip": chip,
                       "s2_paths": [os.path.join(self.s2_path, chip, band + ".png") for band in self.s2_bands]}
                if old_index is not None:
                    rec["old_naip"] = old_index[chip][0]
                records.append(rec)
        # the subset draw uses the global `random` state exactly like the reference (s2-naip_dataset.py:101-103)
        if "train_samples" in opt and train:
            records = random.sample(records, opt["train_samples"])
        self.datapoints = records
        self.data_len = len(records)

    # ------------------------------------------------------------------ storage access
    def _array(self, entry):
        n = int(np.prod(entry["shape"]))
        return torch.from_numpy(np.array(self._bin[entry["off"]:entry["off"] + n]).reshape(entry["shape"]))

    def _hr(self, rec):
        return self._array(rec["hr"]) if self.shard_path else _read_png(rec["naip"])

    def _old_hr(self, rec):
        return self._array(rec["old_hr"]) if self.shard_path else _read_png(rec["old_naip"])

    def _s2_tensor(self, rec):
        """[T_all, C_total, 32, 32] or raises, mirroring s2-naip_dataset.py:180-198"""
        s2_tensor = None
        n_items = len(rec["bands"]) if self.shard_path else len(rec["s2_paths"])
        for i in range(n_items):
            if self.shard_path:
                b = rec["bands"][i]
                if b.get("broken"):
                    raise RuntimeError("undecodable Sentinel-2 PNG")
                if b.get("missing"):
                    s2_img = torch.zeros((self.n_s2_images, 3 if b["tci"] else 1, 32, 32), dtype=torch.uint8)
                else:
                    s2_img = self._array(b)
            else:
                p = rec["s2_paths"][i]
                if not os.path.exists(p):
                    s2_img = torch.zeros((self.n_s2_images, 3 if "tci" in p else 1, 32, 32), dtype=torch.uint8)
                else:
                    s2_img = _read_png(p)
                    s2_img = torch.reshape(s2_img, (s2_img.shape[0], -1, 32, 32)).permute(1, 0, 2, 3)
            s2_tensor = s2_img if i == 0 else torch.cat((s2_tensor, s2_img), dim=1)
        return s2_tensor

    # ------------------------------------------------------------------ reference surface
    def get_tile_weight_sampler(self, tile_weights):
        """s2-naip_dataset.py:132-150"""
        weights = []
        for rec in self.datapoints:
            chip = rec["naip"].split("/")[-1][:-4]
            weights.append(tile_weights[chip] if chip in tile_weights else 1)
        return CustomWeightedRandomSampler(weights, len(self.datapoints))

    def _from_shard(self, rec, index):
        """The sample of `__getitem__` assembled from the index metadata and only the chosen frames' bytes; None = the
        reference would have skipped this datapoint.  Same decisions and the same `random` calls as the generic path."""
        if rec["hr_black"]:
            return None
        k = self.n_s2_images
        T = None
        for b in rec["bands"]:
            if b.get("broken"):
                return None                                    # undecodable PNG -> exception -> skip (:199-201)
            t_b = k if b.get("missing") else b["shape"][0]
            if T is not None and t_b != T:
                return None                                    # torch.cat of bands with different frame counts raises -> skip
            T = t_b
        if T < k:
            return None
        first = rec["bands"][0]
        black = [True] * T if first.get("missing") else first["black"]
        clean = [t for t in range(T) if not black[t]]
        dirty = [t for t in range(T) if black[t]]
        chosen = random.sample(clean, k) if len(clean) >= k else clean + random.sample(dirty, k - len(clean))
        parts = []
        for b in rec["bands"]:
            if b.get("missing"):
                parts.append(torch.zeros((k, 3 if b["tci"] else 1, 32, 32), dtype=torch.uint8))
            else:
                n = int(np.prod(b["shape"]))
                arr = self._bin[b["off"]:b["off"] + n].reshape(b["shape"])
                parts.append(torch.from_numpy(arr[chosen]))    # fancy index on the memmap: copies just these frames
        frames = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        hr = self._array(rec["hr"])
        if self.rand_crop:
            frames, hr = self._augment(frames, hr)
        if not self.use_3d:
            frames = torch.reshape(frames, (-1, 32, 32))
        sample = {"hr": hr, "lr": frames, "Index": index, "Phase": self.split, "Chip": rec["chip"]}
        if self.old_naip_path is not None:
            sample["old_hr"] = self._array(rec["old_hr"])
        return sample

    def _pick_frames(self, stack):
        """frame choice of s2-naip_dataset.py:206-222: frames whose TCI has a black pixel are used only to fill up"""
        clean, dirty = [], []
        for t, frame in enumerate(stack[:, :3]):
            (dirty if has_black_pixels(frame) else clean).append(t)
        k = self.n_s2_images
        chosen = random.sample(clean, k) if len(clean) >= k else clean + random.sample(dirty, k - len(clean))
        return stack[torch.as_tensor(chosen)]

    def _augment(self, frames, hr):
        """random crop to an edge in [24, 32] (LR) / 4x that (HR), resized back with nearest (s2-naip_dataset.py:226-233)"""
        edge = random.randint(24, 32)
        frames = F.interpolate(frames[:, :, :edge, :edge], (32, 32))
        hr = F.interpolate(hr[:, :4 * edge, :4 * edge].unsqueeze(0), (128, 128)).squeeze(0)
        return frames, hr

    def __getitem__(self, index):
        skipped = 0   # rejected samples advance the index cumulatively, as in the reference's while-loop (:158-163)
        while True:
            index += skipped
            if index >= self.data_len:
                index = 0
            rec = self.datapoints[index]
            if self.shard_path:
                sample = self._from_shard(rec, index)
                if sample is None:
                    skipped += 1
                    continue
                return sample
            hr = self._hr(rec)
            ok = not has_black_pixels(hr)                      # partially invalid NAIP chip (:171-175)
            if ok:
                try:
                    stack = self._s2_tensor(rec)                # rare undecodable / inconsistent band files (:180-201)
                    ok = stack.shape[0] >= self.n_s2_images     # too few Sentinel-2 frames (:203-205)
                except Exception:
                    ok = False
            if not ok:
                skipped += 1
                continue
            frames = self._pick_frames(stack)
            if self.rand_crop:
                frames, hr = self._augment(frames, hr)
            if not self.use_3d:
                frames = torch.reshape(frames, (-1, 32, 32))    # channel = t * C + c
            sample = {"hr": hr, "lr": frames, "Index": index, "Phase": self.split, "Chip": rec["chip"]}
            if self.old_naip_path is not None:
                sample["old_hr"] = self._old_hr(rec)
            return sample

    def __len__(self):
        return self.data_len


_register(DATASET_REGISTRY, S2NAIPShardDataset)


class PinnedBatcher:
    """Collates samples straight into two rotating pinned uint8 staging buffers (`lr` [B, T*C, 32, 32], `hr` [B, 3, 128, 128]) --
    the layout `ESRGANTrainer.feed_data` copies to the device with one H2D each."""

    def __init__(self, dataset, batch_size, pin=None):
        self.ds, self.B = dataset, batch_size
        s = dataset[0]
        pin = torch.cuda.is_available() if pin is None else pin
        mk = lambda shape: torch.empty((batch_size,) + tuple(shape), dtype=torch.uint8, pin_memory=pin)
        self.bufs = [{"lr": mk(s["lr"].shape), "hr": mk(s["hr"].shape)} for _ in range(2)]
        self.turn = 0

    def batch(self, indices):
        buf = self.bufs[self.turn]
        self.turn ^= 1
        for i, idx in enumerate(indices):
            s = self.ds[idx]
            buf["lr"][i].copy_(s["lr"])
            buf["hr"][i].copy_(s["hr"])
        return buf

    def batches(self, sampler):
        """pinned batches in the order `sampler` yields indices (a trailing partial batch is dropped: `drop_last` of the train loader)"""
        idx = []
        for i in sampler:
            idx.append(int(i))
            if len(idx) == self.B:
                yield self.batch(idx)
                idx = []


def build_train_sampler(dataset, dataset_opt):
    """The `tile_weights` key of the train dataset options (esrgan_s2naip_urban.yml:25: a JSON {naip chip: weight}) is read by nothing in
    the reference -- ssr/train.py:94-95 only carries a TODO for it, S2NAIPDataset.get_tile_weight_sampler (s2-naip_dataset.py:132-150)
    is never called.  This is the missing wire: the weighted sampler when `tile_weights` is set (a path or an already-loaded dict),
    None (the loader's own sampler) otherwise.  `use_shuffle` must be off with it (yml:27)."""
    tw = dataset_opt.get("tile_weights")
    if not tw:
        return None
    if dataset_opt.get("use_shuffle"):
        raise ValueError("tile_weights: use_shuffle must be False when the weighted tile sampler is used (esrgan_s2naip_urban.yml:27)")
    if not isinstance(tw, dict):
        with open(tw) as fh:
            tw = json.load(fh)
    return dataset.get_tile_weight_sampler(tw)
