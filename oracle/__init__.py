"""Oracle: fp32 pure-PyTorch restatement of the EdgeStyle per-step denoise hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``edgestyle_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of
``bench.py`` do, and there only as the checker / the reported CPU baseline.

PARITY UNPINNED: the reference ships no unit tests, golden vectors or
fixtures for this path, and its arithmetic lives in the third-party ``diffusers==0.26.3``
(pinned at /root/reference/requirements-jetson.txt:25), which is neither vendored under
/root/reference nor installed here.  What IS pinned:

* the pure-torch slices of the reference that execute without diffusers (``ControlNetBlock``,
  ``interleave_tensors``; model/edgestyle_multicontrolnet.py:23-63,479-514) were run in the
  build container by ``tests/golden/make_golden.py`` and their outputs are committed under
  ``tests/golden/`` -- ``oracle.merge`` is checked against them bit-for-bit-tolerance; likewise
  ``VAEControlNetConditioningEmbedding`` + ``_tie_weights`` (model/controllora.py:28-56, run by
  ``tests/golden/make_golden_vae_cond.py`` around the oracle's VAE) pin ``oracle.controllora``'s embedder: zeroed
  ``conv_vae_out`` aliasing ``conv_in``, re-tied to the UNet's conv_in, global-RNG sample x 0.18215; and
  ``EdgeStyleMultiControlNetModel.forward`` (model/edgestyle_multicontrolnet.py:116-171) was run on stub nets by
  ``tests/golden/make_golden_multi_forward.py`` and pins ``oracle.merge.EdgeStyleMultiControlNetModel.forward``;
  ``CachedControlNetModel.forward`` (model/controllora.py:59-287) was run over the oracle's sub-modules by
  ``tests/golden/make_golden_controlnet_forward.py`` and pins ``oracle.sd15.ControlNetModel.forward``;
* the published parameter counts (UNet 859 520 964, ControlNet 361 279 120) and the residual
  shape table (model/edgestyle_onnx_pipeline.py:244-258);
* algebraic identities (LoRA fuse == unfused, zero zero-convs => cond-independent UNet, DDIM
  schedule constants).

``oracle.vae`` (AutoencoderKL: the per-call VAE stages either side of the loop) is likewise a
restatement of diffusers 0.26.3 with no golden vectors in the reference: parity unpinned; pinned on the published SD1.5
VAE parameter count (83 653 863) and the diffusers ``vae/`` state-dict key names.

Everything else follows SURVEY.md Appendix A (diffusers 0.26.3 semantics for SD1.5).
"""
