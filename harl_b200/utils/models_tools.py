"""Device / schedule helpers with the reference's names (harl/utils/models_tools.py)."""
import torch


def init_device(args):
    """models_tools.py:8-25.  There is no CPU fallback for the compute path: ``cuda: False`` (or no GPU)
    yields a CPU device on which only the host-side logic (buffers, configs) works."""
    if args["cuda"] and torch.cuda.is_available():
        device = torch.device("cuda", int(args.get("device_index", torch.cuda.current_device())))
    else:
        device = torch.device("cpu")
    torch.set_num_threads(args["torch_threads"])
    return device


def linear_schedule_lr(epoch, total_num_epochs, initial_lr):
    """models_tools.py:77-87: lr0 * (1 - (epoch - 1) / total)."""
    return initial_lr - (initial_lr * ((epoch - 1) / float(total_num_epochs)))
