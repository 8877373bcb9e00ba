"""CPU oracle for the TAP-CLIP hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline — never as a fallback for the CUDA path.

What it restates (all fp32, torch CPU ops):

* ``clip_standin``   – an open_clip-shaped CLIP (vision tower, text transformer,
  token embedding, text projection) built from stock ``torch.nn`` modules,
  the ``CLIPWrapper`` of ``models/clip_wrapper.py:9-65`` minus its two
  ``open_clip`` calls, a synthetic tokenizer and a seeded weight generator.
* ``tapclip_oracle`` – ``PromptLearner`` (``models/prompt_learner.py:5-70``),
  ``AttributionMonitor`` (``models/attribution_monitor.py:7-36``),
  ``PromptAdjustor`` (``models/prompt_adjustor.py:6-47``) and
  ``FullModel.forward`` (``models/model_wrapper.py:28-100``), both as written
  (the B x n_cls double loop) and de-duplicated.

Parity pinning
--------------
The reference ships no tests, fixtures or golden vectors, and the package that
holds its arithmetic (``open_clip``, un-pinned, un-vendored) is absent here.

* Everything the reference itself owns (FullModel / PromptLearner /
  AttributionMonitor / PromptAdjustor) is PINNED: ``oracle/make_goldens.py``
  imports those modules unmodified from ``/root/reference`` in the build
  container, drives them with the stand-in wrapper, checks the restatement
  against them bit for bit, and commits their outputs under ``tests/golden/``.
* The open_clip model itself is "parity unpinned": it is restated from
  open_clip's published structure and cross-checked against the independent
  HuggingFace ``transformers`` CLIP implementation (``tests/test_oracle_hf.py``).
* Image-side CLS-row attention attribution (north-star extension, not in the
  reference) is "parity unpinned"; it is cross-checked against HF
  ``output_attentions``.
"""
