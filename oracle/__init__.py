"""CPU oracle for the zonal segmentation hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy + torch fp32, eager), the algorithm of
kezakool/flair-for-aigle's zonal inference path so that the CUDA implementation in
``flair_for_aigle_b200`` can be checked against it.  It must never be imported by the
product path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.

PARITY UNPINNED (by the reference): the reference ships no tests, golden vectors or
recorded outputs for this path, and it cannot be imported in
this image (segmentation_models_pytorch, timm, rasterio, geopandas are absent).  The
oracle is therefore pinned by
  * the known-answer tile counts / offsets derived from the reference formulas
     for the grid and window arithmetic (``oracle/grid.py``), and
  * independent in-image implementations of the same third-party arithmetic
    (torchvision ``resnet34``, HF ``ConvNextV2Model``, ``torch.nn.functional``) for the
    model restatements (``oracle/models.py``), see ``tests/test_oracle_models.py``.

Modules
  grid.py      slicing.py:51-112 tile grid, inference.py:300-343 crop/write windows
  convert.py   postprocess.py:9-30 ``convert`` and the intended accumulate/argmax of
               inference.py:468-572
  models.py    smp==0.4.0 / timm model restatements + FLAIR_HUB_Model wiring
  pipeline.py  dataset.py:89-124,174-209 + inference.py:254-355 driven on in-memory rasters
"""
