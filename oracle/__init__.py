"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's detection-forward path.

Nothing under `oracle/` is imported by the product package `dma_yolo_b200/`.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may use it, and
only as the checker or the timed CPU baseline — never as the thing shipped.

Parity status: PINNED.  The reference ships no tests or golden vectors for this path (SURVEY.md
F11), so the restatement is pinned against outputs of the reference itself, executed in the build
container through `oracle/refshim.py` (fixtures under tests/golden/, generator
`oracle/make_golden.py`), and — for NMS, whose arithmetic lives in the third-party dependency
torchvision (requirements.txt:12 `torchvision>=0.8.1`, unpinned; installed 0.26.0) — against the
installed `torchvision.ops.nms` CPU operator on every CPU test run.
"""
