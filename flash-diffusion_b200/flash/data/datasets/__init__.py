"""reference: src/flash/data/datasets/{dataset,datasets_config,collation_fn}.py — `DataModule`, `DataModuleConfig`,
`DataPipeline` with the same constructor arguments.  webdataset is replaced by a small tar-shard reader: members
`<key>.<ext>` of a shard are grouped by key into one sample dict {ext: decoded value} ("jpg"/"png" -> PIL image with
decoder "pil", "json" -> dict, "txt" -> str), shards are split across ranks (`wds.split_by_node`) and DataLoader
workers, samples go through the filter / mapper chain and are collated by key."""
import io
import json
import os
import random
import tarfile
from typing import Callable, List, Optional, Union

import torch
from pydantic.dataclasses import dataclass

from ...config import BaseConfig
from ..filters import BaseFilter, FilterWrapper
from ..mappers import BaseMapper, MapperWrapper


def warn_and_continue(exn):
    import logging
    logging.warning(f"data pipeline: {exn!r}")
    return True


@dataclass
class DataModuleConfig(BaseConfig):
    shards_path_or_urls: Union[str, List[str]] = None
    per_worker_batch_size: int = 16
    num_workers: int = 1
    shuffle_before_split_by_node_buffer_size: Optional[int] = 100
    shuffle_before_split_by_workers_buffer_size: Optional[int] = 100
    shuffle_before_filter_mappers_buffer_size: Optional[int] = 1000
    shuffle_after_filter_mappers_buffer_size: Optional[int] = 1000
    decoder: str = "pil"
    handler: Callable = warn_and_continue
    rename_files_fn: Optional[Callable[[str], str]] = None

    def __post_init__(self):
        super().__post_init__()
        if self.rename_files_fn is not None:
            assert callable(self.rename_files_fn), "rename_files must be a callable"


def custom_collation_fn(samples, combine_tensors=True, combine_scalars=True):
    """reference collation_fn.py:7-41 — the keys COMMON to all samples; per key the type of the first value decides:
    int / float (bools included) -> one numpy array, torch tensors -> `torch.stack`, numpy arrays -> one numpy array,
    anything else -> the plain list.  A scalar / tensor column whose `combine_*` flag is off is dropped, as upstream."""
    import numpy as np
    keys = set.intersection(*[set(s.keys()) for s in samples])
    out = {}
    for k in keys:
        vals = [s[k] for s in samples]
        first = vals[0]
        if isinstance(first, (int, float)):
            if combine_scalars:
                out[k] = np.array(vals)
        elif isinstance(first, torch.Tensor):
            if combine_tensors:
                out[k] = torch.stack(vals)
        elif isinstance(first, np.ndarray):
            if combine_tensors:
                out[k] = np.array(vals)
        else:
            out[k] = vals
    return out


def _decode(ext, data, decoder):
    if ext in ("jpg", "jpeg", "png", "webp") and decoder == "pil":
        from PIL import Image
        return Image.open(io.BytesIO(data)).convert("RGB")
    if ext == "json":
        return json.loads(data)
    if ext in ("txt", "text", "caption"):
        return data.decode("utf-8")
    return data


def _shard_path(url):
    url = url.strip()
    for prefix in ("pipe:cat ", "pipe:", "file:"):
        if url.startswith(prefix):
            url = url[len(prefix):].strip()
    return url


class _ShardDataset(torch.utils.data.IterableDataset):
    def __init__(self, cfg: DataModuleConfig, chain):
        self.cfg, self.chain = cfg, chain

    def _shards(self):
        urls = self.cfg.shards_path_or_urls
        urls = [urls] if isinstance(urls, str) else list(urls)
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        urls = urls[rank::world] if len(urls) >= world else urls          # wds.split_by_node
        info = torch.utils.data.get_worker_info()
        if info is not None and len(urls) >= info.num_workers:
            urls = urls[info.id::info.num_workers]                        # wds.split_by_worker
        return urls

    def _samples(self):
        for url in self._shards():
            try:
                with tarfile.open(_shard_path(url)) as tf:
                    cur_key, cur = None, {}
                    for m in tf:
                        if not m.isfile():
                            continue
                        name = self.cfg.rename_files_fn(m.name) if self.cfg.rename_files_fn else m.name
                        key, _, ext = os.path.basename(name).partition(".")
                        key = os.path.join(os.path.dirname(name), key)
                        if key != cur_key and cur:
                            yield cur
                            cur = {}
                        cur_key = key
                        cur["__key__"] = key
                        cur[ext.lower()] = _decode(ext.lower(), tf.extractfile(m).read(), self.cfg.decoder)
                    if cur:
                        yield cur
            except Exception as e:          # noqa: BLE001
                if not self.cfg.handler(e):
                    raise

    def __iter__(self):
        buf, size = [], self.cfg.shuffle_after_filter_mappers_buffer_size or 0
        for s in self._samples():
            try:
                keep = True
                for step in self.chain:
                    if isinstance(step, (BaseFilter, FilterWrapper)):
                        if not step(s):
                            keep = False
                            break
                    else:
                        s = step(s)
                if not keep:
                    continue
            except Exception as e:          # noqa: BLE001
                if not self.cfg.handler(e):
                    raise
                continue
            s.pop("__key__", None)
            if size > 1:
                buf.append(s)
                if len(buf) >= size:
                    yield buf.pop(random.randrange(len(buf)))
            else:
                yield s
        random.shuffle(buf)
        yield from buf


class DataPipeline:
    """one configuration -> one DataLoader (reference dataset.py:13-145)"""

    def __init__(self, config: DataModuleConfig, filters_mappers=None, batched_filters_mappers=None):
        self.config = config
        self.filters_mappers = list(filters_mappers or [])
        self.batched_filters_mappers = list(batched_filters_mappers or [])
        self.dataset = None

    def setup(self):
        self.dataset = _ShardDataset(self.config, self.filters_mappers)

    def _collate(self, samples):
        batch = custom_collation_fn(samples)
        for step in self.batched_filters_mappers:
            if isinstance(step, (BaseMapper, MapperWrapper)):
                batch = step(batch)
        return batch

    def dataloader(self):
        return torch.utils.data.DataLoader(self.dataset, batch_size=self.config.per_worker_batch_size,
                                           num_workers=self.config.num_workers, collate_fn=self._collate,
                                           drop_last=True)


class DataModule:
    """reference dataset.py:148-208 (`pl.LightningDataModule` there; the Trainer here only needs the two loaders)"""

    def __init__(self, train_config: DataModuleConfig, train_filters_mappers=None, train_batched_filters_mappers=None,
                 eval_config: DataModuleConfig = None, eval_filters_mappers=None, eval_batched_filters_mappers=None):
        self.train_config, self.eval_config = train_config, eval_config
        self.train_pipeline = DataPipeline(train_config, train_filters_mappers, train_batched_filters_mappers)
        self.eval_pipeline = (DataPipeline(eval_config, eval_filters_mappers, eval_batched_filters_mappers)
                              if eval_config is not None else None)

    def setup(self, stage=None):
        self.train_pipeline.setup()
        if self.eval_pipeline is not None:
            self.eval_pipeline.setup()

    def train_dataloader(self):
        return self.train_pipeline.dataloader()

    def val_dataloader(self):
        return self.eval_pipeline.dataloader() if self.eval_pipeline is not None else None


__all__ = ["DataModule", "DataModuleConfig", "DataPipeline", "custom_collation_fn"]
