"""Checkpoint ("workspace") I/O in the reference's file layout (``utils/utils.py:55-148``): a zip named
``<file_name><id>.zip`` holding ``torchModel<id>.pth`` = ``{'epoch', 'state_dict', 'optimizer'}`` and ``args.json``.

``state_dict`` is the model's (same 873 keys / shapes as the reference, so either implementation loads the other's file).
The reference's optimizer is ``Adam(model.parameters())`` -- one state entry per parameter, in ``model.parameters()``
order (``main.py:78``) -- whereas training here runs Adam on ONE flat parameter (``model.flat_parameter_for_optimizer()``).
``flat_optimizer_state_to_reference`` / ``load_flat_optimizer_state`` convert between the two, so a run can be resumed on
either side of the switch.  Pure host-side code (no CUDA needed).
"""
import copy
import json
import os
import zipfile
from zipfile import ZipFile

import torch

PARAM_BLACKLIST = ['epoch_start', 'epochs', 'run_dir', 'ckpt_dir', 'pred_dir']      # utils/utils.py:20
_ADAM_KEYS = ("exp_avg", "exp_avg_sq", "max_exp_avg_sq")


def _param_slices(model):
    """(offset, numel, shape) of every nn.Parameter inside the flat buffer, in ``model.parameters()`` order."""
    by_name = {name: (off, numel, shape) for name, off, numel, shape in model._table}
    return [by_name[name] for name, _ in model.named_parameters()]


def _is_flat_optimizer(model, optimizer):
    ps = [p for g in optimizer.param_groups for p in g["params"]]
    return len(ps) == 1 and ps[0].numel() == model.flat_parameters().numel()


def flat_optimizer_state_to_reference(model, optimizer):
    """``optimizer.state_dict()`` of an Adam built on the flat parameter, re-expressed as the state dict of
    ``Adam(model.parameters())``: per-parameter ``step / exp_avg / exp_avg_sq [/ max_exp_avg_sq]``, ids 0..n-1."""
    sd = optimizer.state_dict()
    flat_state = sd["state"].get(0, {})
    slices = _param_slices(model)
    state = {}
    if flat_state:
        for i, (off, numel, shape) in enumerate(slices):
            ent = {"step": flat_state["step"].clone() if torch.is_tensor(flat_state["step"]) else flat_state["step"]}
            for k in _ADAM_KEYS:
                if k in flat_state:
                    ent[k] = flat_state[k][off:off + numel].detach().reshape(shape).clone().cpu()
            state[i] = ent
    group = {k: copy.deepcopy(v) for k, v in sd["param_groups"][0].items() if k != "params"}
    group["params"] = list(range(len(slices)))
    # the flat optimizer itself always runs with weight_decay = 0 (see load_flat_optimizer_state); the file carries the
    # decay the trainer applies to the trainable entries (main.py:78)
    group["weight_decay"] = float(getattr(optimizer, "reference_weight_decay", group.get("weight_decay", 0.0)))
    return {"state": state, "param_groups": [group]}


def load_flat_optimizer_state(model, optimizer, reference_state_dict):
    """Inverse: scatter a reference-layout Adam state (one entry per parameter) into the flat-parameter optimizer; the
    non-trainable entries of the flat buffer keep zero moments.  Hyper-parameters (lr, betas, ...) follow the file,
    EXCEPT the weight decay: the flat parameter also holds non-trainable buffers (permutations, masks, BatchNorm running
    statistics), and Adam's normalisation would turn ``wd * p`` on their zero gradient into a step of about ``lr``, so the
    flat optimizer is forced to ``weight_decay = 0``.  The file's value (1e-8 for reference checkpoints, main.py:78) is
    kept as ``optimizer.reference_weight_decay`` and returned: pass it to ``train.train_block(weight_decay=...)``, which
    applies it to the trainable entries only.  Returns ``(optimizer, weight_decay)``."""
    flat = model.flat_parameter_for_optimizer()
    slices = _param_slices(model)
    ref_state = reference_state_dict["state"]
    new_state = {}
    if ref_state:
        assert len(ref_state) == len(slices), "optimizer state has %d entries, the model has %d parameters" % (len(ref_state), len(slices))
        first = ref_state[min(ref_state)] if not isinstance(next(iter(ref_state)), str) else ref_state[next(iter(ref_state))]
        ent = {"step": first["step"].clone() if torch.is_tensor(first["step"]) else torch.tensor(float(first["step"]))}
        for k in _ADAM_KEYS:
            if k in first:
                buf = torch.zeros_like(flat.detach())
                for i, (off, numel, shape) in enumerate(slices):
                    src = ref_state[i] if i in ref_state else ref_state[str(i)]
                    buf[off:off + numel] = src[k].reshape(-1).to(buf.device, buf.dtype)
                ent[k] = buf
        new_state[0] = ent
    group = {k: copy.deepcopy(v) for k, v in reference_state_dict["param_groups"][0].items() if k != "params"}
    group["params"] = [0]
    file_wd = float(group.get("weight_decay", 0.0) or 0.0)
    group["weight_decay"] = 0.0
    # keys a newer torch expects but an old file lacks keep the optimizer's current values
    cur = optimizer.state_dict()["param_groups"][0]
    for k, v in cur.items():
        group.setdefault(k, v)
    optimizer.load_state_dict({"state": new_state, "param_groups": [group]})
    optimizer.reference_weight_decay = file_wd
    return optimizer, file_wd


def saveWorkspace(args, model, optimizer, file_name="nsWorkspace", file_id=0):
    """``utils/utils.py:55-93``: model / optimizer state and the program arguments into ``<ckpt_dir>/<file_name><id>.zip``.
    A flat-parameter optimizer is stored in the reference's per-parameter layout."""
    core = getattr(model, "module", model)
    opt_sd = flat_optimizer_state_to_reference(core, optimizer) if _is_flat_optimizer(core, optimizer) else optimizer.state_dict()
    model_file_name = os.path.join(args.ckpt_dir, 'torchModel{:d}.pth'.format(file_id))
    state = {'epoch': file_id, 'state_dict': {k: v.detach().cpu() for k, v in core.state_dict().items()}, 'optimizer': opt_sd}
    torch.save(state, model_file_name)
    args_file_name = os.path.join(args.ckpt_dir, "args.json")
    args_dict = copy.deepcopy({k: v for k, v in vars(args).items() if k != 'device'})
    with open(args_file_name, 'w') as args_file:
        json.dump(args_dict, args_file, indent=4, default=str)
    zip_file_name = os.path.join(args.ckpt_dir, file_name + '{:d}.zip'.format(file_id))
    with ZipFile(zip_file_name, 'w', compression=zipfile.ZIP_DEFLATED) as zipObj:
        zipObj.write(model_file_name, os.path.basename(model_file_name))
        zipObj.write(args_file_name, os.path.basename(args_file_name))
    os.remove(model_file_name)
    os.remove(args_file_name)
    return zip_file_name


def loadWorkspace(args, file_dir, file_name="nsWorkspace", file_id=0, trusted_pickle=False):
    """``utils/utils.py:95-148``: returns ``(args, model_state_dict, optimizer_state_dict)`` or ``None`` when the zip does
    not exist; arguments in the file overwrite ``args`` except the black-listed run-control ones.  The payload holds only
    tensors, dicts, lists and numbers, so it is read with ``weights_only=True`` (no arbitrary code from an untrusted file);
    ``trusted_pickle=True`` opts into full unpickling for files that carry other objects."""
    path = os.path.join(file_dir, file_name + "{:d}.zip".format(file_id))
    if not os.path.exists(path):
        return None
    with ZipFile(path) as zipObj:
        names = zipObj.namelist()
        if "args.json" in names:
            loaded_json = json.loads(zipObj.read("args.json").decode("utf-8"))
            for x in loaded_json:
                if x not in PARAM_BLACKLIST:
                    setattr(args, x, loaded_json[x])
        torch_name = 'torchModel{:d}.pth'.format(file_id)
        with zipObj.open(torch_name) as f:
            import io
            param_dict = torch.load(io.BytesIO(f.read()), map_location="cpu", weights_only=not trusted_pickle)
    return args, param_dict['state_dict'], param_dict['optimizer']
