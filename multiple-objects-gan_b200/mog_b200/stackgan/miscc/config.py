"""Global ``cfg`` of the COCO StackGAN program: keys and defaults of
``code/coco/stackgan/miscc/config.py:8-58`` (+ the ``MOG`` section of this implementation)."""
from ..._cfg_util import edict, make_cfg


def _defaults():
    c = edict()
    c.DATASET_NAME = 'coco'
    c.EMBEDDING_TYPE = 'cnn-rnn'
    c.CONFIG_NAME = ''
    c.GPU_ID = '0'
    c.CUDA = True
    c.WORKERS = 6
    c.NET_G = ''
    c.NET_D = ''
    c.STAGE1_G = ''
    c.DATA_DIR = ''
    c.IMG_DIR = ''
    c.VIS_COUNT = 64
    c.Z_DIM = 100
    c.IMSIZE = 64
    c.STAGE = 1
    c.USE_LOCAL_PATHWAY = True
    c.USE_BBOX_LAYOUT = True
    c.TRAIN = edict(FLAG=True, BATCH_SIZE=64, MAX_EPOCH=600, SNAPSHOT_INTERVAL=50, PRETRAINED_MODEL='',
                    PRETRAINED_EPOCH=600, LR_DECAY_EPOCH=600, DISCRIMINATOR_LR=2e-4, GENERATOR_LR=2e-4,
                    COEFF=edict(KL=2.0))
    c.GAN = edict(CONDITION_DIM=128, DF_DIM=64, GF_DIM=128, R_NUM=4)
    c.TEXT = edict(DIMENSION=1024)
    c.MOG = edict(ALIGN_CORNERS=False)
    return c


cfg, cfg_from_file, reset_cfg = make_cfg(_defaults)
