import os


def set_cuda_visible_devices(gpu_ids=None, **kwargs):
    """train.py:58 / test.py:65: restrict the process to the given GPUs ("0,1")."""
    if gpu_ids is not None:
        os.environ["CUDA_VISIBLE_DEVICES"] = str(gpu_ids)
