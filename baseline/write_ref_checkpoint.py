"""Writes a checkpoint with the UNMODIFIED reference trainer's ``save_model`` (code/coco/attngan/trainer.py:173-199) after one
reference training step at a tiny configuration -- test infrastructure (tests/test_host_cpu.py::test_reference_checkpoint_loads).

    python baseline/write_ref_checkpoint.py <model_dir> <epoch>
"""
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
warnings.filterwarnings("ignore")
from baseline import ref_harness as H  # noqa: E402

if __name__ == "__main__":
    model_dir, epoch = sys.argv[1], int(sys.argv[2])
    ns = H.load("attngan", "cpu")
    from mog_b200 import synth
    c = dict(GF_DIM=4, DF_DIM=4, Z_DIM=20, R_NUM=1, EMBEDDING_DIM=16, T=6)
    st = H.AttnGANStep(ns, 2, c=c, seed=5, device="cpu", damsm=True, encoder=synth.StandInEncoder(16), init="fill", logit_scale=0.02)
    st.step()
    import importlib
    T = importlib.import_module("trainer")          # the reference's trainer.py
    fake_self = types.SimpleNamespace(model_dir=model_dir)
    T.condGANTrainer.save_model(fake_self, st.netG, st.avg_param_G, st.netsD, st.optimizerG, st.optimizersD, epoch)
    print("written")
