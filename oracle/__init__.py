"""oracle/ — TEST INFRASTRUCTURE ONLY.

A CPU restatement (PyTorch fp32 on CPU + numpy) of the reference's algorithms for the
interpretation hot path: pt/mask.py, pt/models/I3D_doubled[_kth].py, pt/models/CLSTM_4.py,
pt/models/convolution_lstm.py, pt/grad_cam_videos.py and the mask loop of
pt/FindMasksComparison_I3D_smth.py (pt/ = /root/reference/video_features_pytorch/).

Nothing in the product package imports this directory; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do, and only as the checker / baseline.

Pinning: the reference ships no tests, fixtures or golden vectors, so the oracle is
pinned against outputs of the UNMODIFIED reference modules imported in the authoring container:
`python oracle/pin_against_reference.py` asserts oracle == reference on seeded inputs and writes
the known-answer vectors to tests/golden/*.npz (committed, with that script).  On the GPU box
/root/reference does not exist; the tests there use this restatement and the committed vectors.
"""
