"""CPU fp32 oracle for the DGP scoremap-and-graph hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deepgraphpose_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the reported CPU baseline -- never as the product path.

What it is: a restatement, in numpy + torch CPU ops, of the reference's TF1
graph for this path (citations are relative to /root/reference):

* ``tf_ops``      -- TF-SAME padding rules, slim ``conv2d_same``, SAME max-pool,
                     TF ``conv2d_transpose`` alignment, frozen batch-norm
                     (third-party tensorflow==1.15 / tf.contrib.slim, not vendored
                     in the reference; call sites
                     src/DeepLabCut/deeplabcut/pose_estimation_tensorflow/nnet/pose_net.py:9-26,50-52).
* ``resnet_v1``   -- slim ``resnet_v1_50`` at output_stride=16, is_training=False
                     (pose_net.py:14-16, 36-54).
* ``pose_net``    -- ``PoseNet.extract_features / test / inference`` and
                     ``prediction_layer`` (pose_net.py:18-163),
                     ``extract_cnn_output`` / ``argmax_pose_predict``
                     (nnet/predict.py:45-77).
* ``dgp_ops``     -- ``argmax_2d_from_cm`` (src/deepgraphpose/models/fitdgp_util.py:281-402),
                     the ``estimate_pose`` per-frame read-out
                     (src/deepgraphpose/models/eval.py:328-357), the
                     ``evaluate_dgp`` 'dgp' locref branch (eval.py:751-785),
                     ``combine_all_marker`` (fitdgp_util.py:232-272).
* ``dgp_loss``    -- the full ``dgp_loss`` graph incl. skeleton and temporal
                     cliques (src/deepgraphpose/models/fitdgp.py:848-1144) and the
                     optimizer step (fitdgp.py:706-713).

Pinning status: the reference holds no golden vectors / known-answer tests for
this path and TensorFlow 1.x cannot be installed in this
image, so the TF *op* semantics are restated from the published TF 1.15
behaviour.  The *composition* of those ops is pinned against the reference's own
source: ``oracle/tf1_shim`` executes the unmodified reference functions
(``argmax_2d_from_cm``, ``dgp_loss``, ``PoseNet`` ...) imported from
/root/reference on top of a small TF1-API emulation, and
``tests/golden/make_golden.py`` commits the resulting vectors under
``tests/golden/``.  See DESIGN.md "Oracle".
"""
