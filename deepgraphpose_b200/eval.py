"""Drop-in for the hot-path part of deepgraphpose.models.eval (reference: src/deepgraphpose/models/eval.py).

``setup_dgp_eval_graph`` :147-214 and ``estimate_pose`` :217-372 keep their names, argument meaning and return
values; the TF graph/session are replaced by the sm_100a engine behind the C ABI.  Video decode (moviepy in the
reference) and the csv/h5 export are outside the path: ``estimate_pose`` takes a video path (decoded
with OpenCV when present) or an in-memory uint8 array of frames.
"""
import os

import numpy as np
import torch

from . import synthetic
from .engine import Engine
from .session import Handle, Session


def _cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def load_variables(dgp_model_file, num_joints, location_refinement):
    """Weights for the graph.  Accepts a ``{tf_var_name: ndarray}`` dict, a ``.npz`` of such arrays, or
    ``"synthetic"`` / ``"synthetic:<seed>"`` (random-init weights, BASELINE.json configs), or the prefix of a TensorFlow
    checkpoint bundle (``snapshot-step2-final--0`` -> ``.index`` + ``.data-0000N-of-0000M``, read by tf_checkpoint.py)."""
    if isinstance(dgp_model_file, dict):
        return dgp_model_file
    name = str(dgp_model_file)
    if name.startswith("synthetic"):
        seed = int(name.split(":")[1]) if ":" in name else 0
        return synthetic.make_weights(num_joints, seed=seed, location_refinement=location_refinement)
    if name.endswith(".npz") and os.path.exists(name):
        with np.load(name) as z:
            return {k: z[k] for k in z.files}
    prefix = name[:-6] if name.endswith(".index") else name
    if os.path.exists(prefix + ".index"):
        # a TensorFlow checkpoint bundle (DLC / DGP snapshot, resnet_v1_50.ckpt): restorer.restore without TensorFlow
        from . import tf_checkpoint
        return tf_checkpoint.model_variables(tf_checkpoint.read_checkpoint(prefix))
    raise FileNotFoundError("no weights at %r: pass a {tf_var_name: ndarray} dict, an .npz of it, 'synthetic[:seed]', or the "
                            "prefix of a TensorFlow checkpoint (<prefix>.index + <prefix>.data-*)" % (name,))


def setup_dgp_eval_graph(dlc_cfg, dgp_model_file, loc_ref=False, gauss_len=1, gamma=1, device=None):
    """eval.py:147-214.  Returns (sess, mu_n, softmax_tensor, scmap, locref, inputs)."""
    nj = int(_cfg_get(dlc_cfg, "num_joints"))
    net_type = _cfg_get(dlc_cfg, "net_type", "resnet_50")
    if net_type != "resnet_50":
        raise ValueError("only net_type='resnet_50' is on the B200 path (got %r)" % (net_type,))
    eng = Engine(nj, location_refinement=bool(loc_ref), device=device,
                 stride=float(_cfg_get(dlc_cfg, "stride", 8.0)),
                 locref_stdev=float(_cfg_get(dlc_cfg, "locref_stdev", 7.2801)),
                 mean_pixel=tuple(_cfg_get(dlc_cfg, "mean_pixel", (123.68, 116.779, 103.939))),
                 precision=_cfg_get(dlc_cfg, "precision", "fp16"))
    eng.load_weights(load_variables(dgp_model_file, nj, bool(loc_ref)))
    handles = {
        "inputs": Handle("Placeholder:0", "inputs"),
        "mu_n": Handle("Sum:0", "mu_n"),
        "softmax_tensor": Handle("truediv:0", "softmax_tensor"),
        "scmap": Handle("pose/part_pred/block4/BiasAdd:0", "scmap"),
        "locref": Handle("pose/locref_pred/block4/BiasAdd:0", "locref") if loc_ref else None,
    }
    sess = Session(eng, handles, gamma=gamma, gauss_len=gauss_len)
    return sess, handles["mu_n"], handles["softmax_tensor"], handles["scmap"], handles["locref"], handles["inputs"]


def _iter_video(video_file):
    import cv2
    cap = cv2.VideoCapture(str(video_file))
    if not cap.isOpened():
        raise IOError("cannot open video %s" % video_file)
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        yield frame[:, :, ::-1]  # BGR -> RGB, as moviepy delivers
    cap.release()


def estimate_pose_frames(engine, frames, batch=16, gamma=1.0, gauss_len=1.0, stride=None):
    """The estimate_pose frame loop (eval.py:306-357) on in-memory frames: uint8 (T,H,W,3), numpy or CPU tensor.

    Host buffers in, host arrays out; the H2D copy of every batch and the D2H copy of its results happen inside.
    Returns dict(x, y, likelihoods, mu_likelihoods, markers) with the reference's shapes.
    """
    ft = frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames))
    mu, peak, lik = engine.estimate_pose_host(ft, batch=batch, gamma=gamma, gauss_len=gauss_len)
    stride = engine.stride if stride is None else stride
    markers = mu.numpy().astype(np.float64)
    xr = markers[:, :, 1] * stride + 0.5 * stride
    yr = markers[:, :, 0] * stride + 0.5 * stride
    return {"x": xr, "y": yr, "likelihoods": lik.numpy().astype(np.float64),
            "mu_likelihoods": peak.numpy().astype("int"), "markers": markers}


def _readout_dict(mu, peak, lik, stride):
    markers = mu.numpy().astype(np.float64)
    return {"x": markers[:, :, 1] * stride + 0.5 * stride, "y": markers[:, :, 0] * stride + 0.5 * stride,
            "likelihoods": lik.numpy().astype(np.float64), "mu_likelihoods": peak.numpy().astype("int"), "markers": markers}


def estimate_pose_stream(engine, source, H, W, n_frames=None, batch=16, gamma=1.0, gauss_len=1.0, chunk=65536):
    """The estimate_pose frame loop (eval.py:306-357) over a frame SOURCE instead of an in-memory array: an iterator of uint8
    (H,W,3) RGB frames (video decoder) or a pinned uint8 tensor (P,H,W,3) cycled for ``n_frames`` frames.  Frames stream
    through a pinned ring of 3 x ``batch`` frames (Engine.estimate_pose_stream): host memory does not grow with the video.
    ``n_frames`` may be None for iterators of unknown length (results are collected ``chunk`` frames at a time)."""
    parts = []
    if isinstance(source, torch.Tensor):
        if n_frames is None:
            n_frames = source.shape[0]
        parts.append(engine.estimate_pose_stream(source, H, W, n_frames, batch, gamma, gauss_len))
    else:
        it = iter(source)
        left = None if n_frames is None else int(n_frames)
        while left is None or left > 0:
            want = chunk if left is None else min(chunk, left)
            mu, peak, lik = engine.estimate_pose_stream(it, H, W, want, batch, gamma, gauss_len)
            parts.append((mu, peak, lik))
            if left is not None:
                left -= mu.shape[0]
            if mu.shape[0] < want:
                break
    mu, peak, lik = (torch.cat([p[i] for p in parts]) for i in range(3))
    return _readout_dict(mu, peak, lik, engine.stride)


def estimate_pose_sharded(engine, source, T, H, W, edges=(), ws=None, ws_max=None, wt_max=0.0, batch=16, gamma=1.0,
                          gauss_len=1.0, group=None, gather=True, timings=None):
    """estimate_pose over a T-frame video sharded contiguously by frame over the ranks of ``group`` (one process
    per GPU): rank r streams frames [a_r, b_r) (``sharding.shard_range``) through its own engine, the ranks exchange the
    soft-argmax of their first frame (one all_gather of nj*8 bytes per rank) so that the temporal potential at every shard
    edge uses the exact neighbour, the skeleton / temporal potentials of the shard are evaluated on the device, and the
    per-frame results are all-gathered back into video order.  The reference has no multi-GPU inference; the per-frame
    arithmetic is that of estimate_pose (eval.py:306-357) and of the clique terms of dgp_loss (fitdgp.py:1063-1083).

    ``source``: a pinned uint8 tensor (P,H,W,3) cycled as the video (frame t = source[t % P]), or a callable
    ``(start, stop) -> iterator of frames`` that opens the rank's own range of a real video.
    Returns the estimate_pose dict plus 'skel' (T,nl), 'temporal' (T,nj; last row 0), 'e_skel', 'e_temp' (T) -- for the
    whole video on every rank when ``gather`` (else for the rank's shard) -- and 'shard' = (a_r, b_r).
    ``timings``: an optional dict that receives the wall-clock seconds of the phases ('stream', 'potentials', 'gather')."""
    import time
    import torch.distributed as dist
    from . import sharding
    on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    if T < world:
        raise ValueError("a %d-frame video cannot be sharded over %d ranks" % (T, world))
    a, b = sharding.shard_range(T, rank, world)
    t_start = time.perf_counter()
    if isinstance(source, torch.Tensor):
        mu, peak, lik = engine.estimate_pose_stream(source, H, W, b - a, batch, gamma, gauss_len, start=a)
    else:
        mu, peak, lik = engine.estimate_pose_stream(source(a, b), H, W, b - a, batch, gamma, gauss_len)
    if mu.shape[0] != b - a:
        raise RuntimeError("rank %d: the source delivered %d of its %d frames" % (rank, mu.shape[0], b - a))
    dev = engine.device
    t_stream = time.perf_counter()
    mu_d = mu.to(dev, non_blocking=True)
    halo = sharding.exchange_halo(mu_d[0].contiguous(), group) if world > 1 else None
    edges = [tuple(e) for e in edges]
    pot = engine.potentials(mu_d, edges, halo_next=halo, ws=ws, ws_max=ws_max, wt_max=wt_max)
    temporal = pot["temporal"]
    if temporal.shape[0] < b - a:       # the video's last frame has no successor
        temporal = torch.cat([temporal, torch.zeros((1, engine.nj), dtype=temporal.dtype, device=dev)])
    local = {"mu": mu_d, "peak": peak.to(dev, non_blocking=True), "lik": lik.to(dev, non_blocking=True),
             "temporal": temporal, "skel": pot["skel"].t().contiguous(), "e_temp": pot["e_temp"]}
    if pot["e_skel"] is not None:
        local["e_skel"] = pot["e_skel"]
    if timings is not None:
        torch.cuda.synchronize(dev)
    t_pot = time.perf_counter()
    full = {k: (sharding.gather_frames(v, T, group) if gather else v) for k, v in local.items()}
    out = _readout_dict(full["mu"].cpu(), full["peak"].cpu(), full["lik"].cpu(), engine.stride)
    for k in ("temporal", "skel", "e_temp", "e_skel"):
        if k in full:
            out[k] = full[k].cpu().numpy()
    out["shard"] = (a, b)
    if timings is not None:
        timings.update(stream=t_stream - t_start, potentials=t_pot - t_stream, gather=time.perf_counter() - t_pot)
    return out


def _video_source(video_file, new_size=None, crop_size=None):
    """(frame iterator, H, W, frame count or None, scale_x, scale_y) for a video file: frames are decoded one at a time
    (OpenCV), resized / cropped the way the reference does per frame (eval.py:309-326), and never held beyond the ring."""
    import cv2
    cap = cv2.VideoCapture(str(video_file))
    if not cap.isOpened():
        raise IOError("cannot open video %s" % video_file)
    w0, h0 = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    count = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    cap.release()
    scale_x = scale_y = 1.0
    H, W = h0, w0
    if new_size is not None:
        scale_x, scale_y = w0 / new_size[1], h0 / new_size[0]
        H, W = int(new_size[0]), int(new_size[1])
    if crop_size is not None:
        W, H = int(crop_size[2] - crop_size[0]), int(crop_size[3] - crop_size[1])

    def frames():
        if new_size is None and crop_size is None:
            yield from _iter_video(video_file)
            return
        from PIL import Image
        for fr in _iter_video(video_file):
            im = Image.fromarray(np.ascontiguousarray(fr))
            if new_size is not None:
                im = im.resize(size=(new_size[1], new_size[0]))
            if crop_size is not None:
                im = im.crop(crop_size)
            yield np.asarray(im)

    return frames(), H, W, (count if count > 0 else None), scale_x, scale_y


def estimate_pose(proj_cfg_file, dgp_model_file, video_file, output_dir, shuffle=1, save_pose=True, save_str="",
                  new_size=None, crop_size=None, batch=16):
    """eval.py:217-372.  ``proj_cfg_file`` may be a dict-like dlc_cfg (num_joints, stride, ...) or a DLC project yaml with
    ``bodyparts``; ``video_file`` a path or a uint8 array (T,H,W,3).  Returns {'x','y','likelihoods'} (T,nj) each.
    A video file is STREAMED: decoded frame by frame (as the reference's ``clip.iter_frames()``, eval.py:306) into a pinned
    ring that the GPU drains batch by batch, so a 100 k-frame clip needs no more host memory than a 100-frame one."""
    if isinstance(proj_cfg_file, (dict,)) or hasattr(proj_cfg_file, "num_joints"):
        dlc_cfg = proj_cfg_file
    else:
        import yaml
        with open(proj_cfg_file, "r") as stream:
            proj = yaml.safe_load(stream)
        dlc_cfg = {"num_joints": len(proj["bodyparts"]), "all_joints_names": list(proj["bodyparts"]), "stride": 8.0}
    save_file = None
    scale_x = scale_y = 1.0
    if isinstance(video_file, (str, os.PathLike)):
        f = os.path.basename(str(video_file)).rsplit(".", 1)
        save_file = os.path.join(str(output_dir), f[0] + "_labeled%s" % save_str)
        if os.path.exists(save_file + ".csv"):
            print("labels already exist! video at %s will not be processed" % video_file)
            return save_file + ".csv"
        source, H, W, n_frames, scale_x, scale_y = _video_source(video_file, new_size, crop_size)
        sess, mu_n, _, scmap, _, inputs = setup_dgp_eval_graph(dlc_cfg, dgp_model_file)
        # the container's frame count is a hint only (it can be off by a few frames): read until the decoder runs dry
        res = estimate_pose_stream(sess.engine, source, H, W, None, batch=batch)
    else:
        frames = np.asarray(video_file)
        if new_size is not None or crop_size is not None:
            from PIL import Image
            out = []
            for fr in frames:
                im = Image.fromarray(fr)
                if new_size is not None:
                    scale_x = im.width / new_size[1]
                    scale_y = im.height / new_size[0]
                    im = im.resize(size=(new_size[1], new_size[0]))
                if crop_size is not None:
                    im = im.crop(crop_size)
                out.append(np.asarray(im))
            frames = np.stack(out)
        sess, mu_n, _, scmap, _, inputs = setup_dgp_eval_graph(dlc_cfg, dgp_model_file)
        res = estimate_pose_frames(sess.engine, frames, batch=batch)
    sess.close()
    labels = {"x": res["x"] * scale_x, "y": res["y"] * scale_y, "likelihoods": res["likelihoods"]}
    if save_pose and save_file is not None:
        os.makedirs(os.path.dirname(save_file) or ".", exist_ok=True)
        names = _cfg_get(dlc_cfg, "all_joints_names", ["joint%d" % i for i in range(labels["x"].shape[1])])
        scorer = os.path.basename(str(dgp_model_file)) if isinstance(dgp_model_file, (str, os.PathLike)) else "dgp_b200"
        export_pose_like_dlc(labels, scorer, list(names), save_file)
    return labels


def export_pose_like_dlc(labels, scorer, joints_names, save_file):
    """eval.py:621-645: the DeepLabCut table layout -- columns (scorer, bodypart, x | y | likelihood), one row per frame --
    written as ``save_file + '.csv'`` (three header rows + frame index column, what DLC / ``plot_dgp`` /
    ``load_pose_from_dlc_to_dict`` read) and, when pytables is installed, as ``save_file + '.h5'`` (key df_with_missing)."""
    import pandas as pd
    x = np.asarray(labels["x"])
    table = np.empty((x.shape[0], 3 * x.shape[1]), dtype=x.dtype)
    table[:, 0::3], table[:, 1::3], table[:, 2::3] = x, labels["y"], labels["likelihoods"]
    columns = pd.MultiIndex.from_product([[scorer], list(joints_names), ["x", "y", "likelihood"]],
                                         names=["scorer", "bodyparts", "coords"])
    frame = pd.DataFrame(table, columns=columns, index=np.arange(x.shape[0]))
    try:
        frame.to_hdf(save_file + ".h5", key="df_with_missing", format="table", mode="w")
    except ImportError:   # pytables is optional here; the csv carries the same table
        pass
    frame.to_csv(save_file + ".csv")


def load_pose_from_dlc_to_dict(filename):
    """eval.py:648-653: read a DLC csv back into {'x', 'y', 'likelihoods'} (T, nj) arrays."""
    body = np.genfromtxt(filename, delimiter=",", skip_header=3)[:, 1:]   # missing values (empty cells) become NaN
    return {"x": body[:, 0::3], "y": body[:, 1::3], "likelihoods": body[:, 2::3]}


def evaluate_dgp_frames(engine, frames, loc_ref=True, loc_ref_calc="dlc", batch=16, gamma=1.0, gauss_len=1.0):
    """The per-image pose read-out of ``evaluate_dgp`` (eval.py:744-790) for a uint8 (T,H,W,3) array: returns (T, nj*3)
    rows of (x, y, likelihood) like ``PredicteData``.  Branches as in the reference: loc_ref with ``loc_ref_calc='dlc'`` ->
    argmax_pose_predict (global peak + locref offset, likelihood = sigmoid at the peak); ``'dgp'`` -> soft-argmax + softmax
    weighted locref offset (likelihood column 1); no loc_ref -> soft-argmax only."""
    frames = np.asarray(frames)
    T = frames.shape[0]
    out = np.empty((T, engine.nj * 3), np.float64)
    for t0 in range(0, T, batch):
        fr = torch.from_numpy(np.ascontiguousarray(frames[t0:t0 + batch])).to(engine.device)
        logits, locref = engine.forward(fr, want_locref=bool(loc_ref))
        if loc_ref and loc_ref_calc.lower() == "dlc":
            pose = engine.softargmax(logits, locref, gamma, gauss_len, want=("dlc_pose",))["dlc_pose"]
        elif loc_ref:
            pose = engine.soft_pose(logits, locref, gamma, gauss_len)
        else:
            mu = engine.softargmax(logits, None, gamma, gauss_len, want=("mu",))["mu"]
            pose = torch.cat([mu.flip(2) * engine.stride + 0.5 * engine.stride, torch.ones_like(mu[:, :, :1])], dim=2)
        out[t0:t0 + fr.shape[0]] = pose.reshape(fr.shape[0], -1).double().cpu().numpy()
    return out


def pairwisedistances(DataCombined, scorer1, scorer2, pcutoff=-1, bodyparts=None):
    """DeepLabCut's evaluate.pairwisedistances (called at eval.py:803-804) on the (scorer, bodypart, coord) column index:
    per frame and bodypart the Euclidean distance between the two scorers' (x, y), and the same masked to predictions of
    ``scorer2`` with likelihood >= pcutoff.  Returns (RMSE, RMSEpcutoff) DataFrames."""
    import pandas as pd
    mask = DataCombined[scorer2].xs("likelihood", level=1, axis=1) >= pcutoff
    if bodyparts is None:
        pointwise = (DataCombined[scorer1] - DataCombined[scorer2]) ** 2
    else:
        pointwise = (DataCombined[scorer1][bodyparts] - DataCombined[scorer2][bodyparts]) ** 2
        mask = mask[bodyparts]
    rmse = np.sqrt(pointwise.xs("x", level=1, axis=1) + pointwise.xs("y", level=1, axis=1))
    return rmse, rmse[mask]


def _read_collected_data(project_path, trainingset_folder, scorer):
    """Human labels of a DLC project: ``CollectedData_<scorer>.h5`` (key df_with_missing) as the reference reads it
    (eval.py:724-725), or the ``.csv`` DLC writes next to it when pytables is not installed."""
    import pandas as pd
    base = os.path.join(project_path, trainingset_folder, "CollectedData_" + scorer)
    try:
        return pd.read_hdf(base + ".h5", "df_with_missing")
    except (ImportError, FileNotFoundError, OSError):
        return pd.read_csv(base + ".csv", header=[0, 1, 2], index_col=0)


def evaluate_dgp(proj_cfg_file, dgp_model_file, shuffle=1, loc_ref=None, loc_ref_calc="dlc", *, dlc_cfg=None, batch=8,
                 device=None):
    """eval.py:656-813: RMSE per labelled frame / joint of a DLC project, train and test errors printed, the RMSE DataFrame
    returned.  ``proj_cfg_file`` is the project's config.yaml; the training-set folder, the train / test split
    (Documentation_data-*.pickle) and the pose config follow DeepLabCut's layout (auxiliaryfunctions.GetTrainingSetFolder /
    GetDataandMetaDataFilenames / LoadMetadata, restated here because deeplabcut is not a dependency).  ``dlc_cfg`` may supply
    the pose config (num_joints, all_joints_names, stride, location_refinement, locref_stdev) instead of the project's
    train/pose_cfg.yaml.  Frames are read with PIL and evaluated in batches of equal size through ``evaluate_dgp_frames``."""
    import pickle
    import pandas as pd
    import yaml
    from PIL import Image
    with open(proj_cfg_file, "r") as stream:
        proj_config = yaml.safe_load(stream)
    project_path = proj_config.get("project_path") or os.path.dirname(os.path.abspath(str(proj_cfg_file)))
    task, date = proj_config["Task"], proj_config["date"]
    iteration = proj_config.get("iteration", 0)
    trainingset_folder = os.path.join("training-datasets", "iteration-" + str(iteration), "UnaugmentedDataSet_" + task + date)
    train_fraction = proj_config["TrainingFraction"][0]
    if dlc_cfg is None:
        model_folder = os.path.join("dlc-models", "iteration-" + str(iteration),
                                    task + date + "-trainset" + str(int(train_fraction * 100)) + "shuffle" + str(shuffle))
        with open(os.path.join(project_path, model_folder, "train", "pose_cfg.yaml"), "r") as stream:
            dlc_cfg = yaml.safe_load(stream)
    dlc_cfg = dict(dlc_cfg) if isinstance(dlc_cfg, dict) else dict(vars(dlc_cfg))
    loc_ref = bool(dlc_cfg.get("location_refinement", True)) if loc_ref is None else bool(loc_ref)
    dlc_cfg["location_refinement"] = loc_ref
    dlc_cfg.setdefault("stride", 8.0)
    sess, mu_n, softmax_tensor, scmap_tf, locref_tf, inputs = setup_dgp_eval_graph(dlc_cfg, dgp_model_file, loc_ref=loc_ref,
                                                                                   device=device)
    Data = _read_collected_data(project_path, trainingset_folder, proj_config["scorer"])
    comparisonbodyparts = list(proj_config["bodyparts"])
    meta = os.path.join(project_path, trainingset_folder,
                        "Documentation_data-" + task + "_" + str(int(train_fraction * 100)) + "shuffle" + str(shuffle) + ".pickle")
    with open(meta, "rb") as f:
        _, trainIndices, testIndices, _ = pickle.load(f)
    names = dlc_cfg.get("all_joints_names") or comparisonbodyparts
    nj = len(names)
    PredicteData = np.ones((len(Data.index), 3 * nj))
    print("Analyzing data...")
    # images of one size go through the engine together (the reference runs them one by one, eval.py:741-742)
    by_shape = {}
    for imageindex, imagename in enumerate(Data.index):
        name = imagename if isinstance(imagename, str) else os.path.join(*imagename)
        image = np.asarray(Image.open(os.path.join(project_path, name)).convert("RGB"))
        by_shape.setdefault(image.shape, []).append((imageindex, image))
    for shape, items in by_shape.items():
        frames = np.stack([im for _, im in items])
        pose = evaluate_dgp_frames(sess.engine, frames, loc_ref, loc_ref_calc, batch=batch)
        for (imageindex, _), row in zip(items, pose):
            PredicteData[imageindex, :] = row
    sess.close()
    DLCscorer = "DGP"
    index = pd.MultiIndex.from_product([[DLCscorer], list(names), ["x", "y", "likelihood"]], names=["scorer", "bodyparts", "coords"])
    DataMachine = pd.DataFrame(PredicteData, columns=index, index=Data.index.values)
    DataCombined = pd.concat([Data.T, DataMachine.T], axis=0).T
    RMSE, _ = pairwisedistances(DataCombined, proj_config["scorer"], DLCscorer, proj_config.get("pcutoff", 0.1), comparisonbodyparts)
    testerror = np.nanmean(RMSE.iloc[testIndices].values.flatten())
    trainerror = np.nanmean(RMSE.iloc[trainIndices].values.flatten())
    print("Train error:", np.round(trainerror, 2), " pixels")
    print("Test error:", np.round(testerror, 2), " pixels")
    return RMSE
