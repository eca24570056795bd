"""`flash.data` — the reference's data-module surface (src/flash/data/{datasets,filters,mappers}), re-implemented without
`webdataset` / `pytorch_lightning` (neither is installable offline; SURVEY.md §2 marks the pipeline itself out of the
hot path): the same class and config names, the same filter / mapper semantics, and a small tar-shard reader that
understands the webdataset layout (`<key>.<ext>` members grouped by key) for the `pipe:cat x.tar` / path URLs the
example scripts pass.  It exists so that `examples/train_flash_*.py` import and run unchanged."""
