"""The ``pointnet2_stack`` ops that SA layers >= 1 and the head's VSA module call: voxel query
(+ dilated), stacked grouping and stacked furthest point sampling -- same names and arguments as
``/root/reference/pcdet/ops/pointnet2/pointnet2_stack`` (``pointnet2_stack_cuda`` = its pybind module re-hosted on the
C ABI, ``pointnet2_utils`` / ``voxel_query_utils`` = its Python wrappers)."""
