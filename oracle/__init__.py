"""CPU oracle for the per-frame face path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (facerec_b200/) never does and fails loudly without its CUDA library.

PARITY UNPINNED for the network arithmetic: the reference delegates RetinaFace and FaceNet to
third-party pip packages (`retinaface`, `keras_facenet`, both unpinned: requirements.txt:6,10;
call sites facerec/detector.py:64,68 and facerec/extract.py:164-167,392) that are neither vendored
under the reference checkout nor installable offline, and the reference holds no tests, golden
vectors or fixtures.  The oracle therefore restates the *published*
architectures BASELINE.json names (RetinaFace MobileNet0.25 / ResNet50 + FPN/SSH + PriorBox decode
+ NMS; FaceNet InceptionResnetV1 + fixed_image_standardization + L2 norm) in plain fp32 PyTorch, and
anchors everything the reference itself fixes on the reference's own call sites: the detect() result
schema, the 0.95 threshold, min-size filter and int() keypoints (facerec/detector.py:66-79), the
crop arithmetic (facerec/extract.py:38-53,157-164), SORT (facerec/sort.py), the distance calls
(facerec/cluster.py:41, facerec/classify_knn.py:212,227: pinned here against scipy / scikit-learn,
the very libraries the reference calls) and the file formats (facerec/utils/utils.py:90-120).
One network piece IS pinned to a canonical implementation available here: the ResNet-50 body of the R50
detector reproduces torchvision.models.resnet50's own layer2/3/4 activations from BN-folded weights
(tests/test_oracle_cpu.py::test_oracle_r50_backbone_matches_torchvision); NMS is pinned to
torchvision.ops.nms and the 160x160 resize to cv2.resize.  MobileNet0.25, FPN/SSH/heads and
InceptionResnetV1 have no installed counterpart and stay unpinned.
"""
