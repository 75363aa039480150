"""Host-side mirror of the node's frame loop, MotionDetectionNode::imageCallback
(ros/src/motion_detection_node.cpp:235-430), without ROS: same parameters, same frame bookkeeping, same log rows.

  skip_frames        node.cpp:237-247   only every skip_frames-th image enters the window
  raw_images_        node.cpp:248-261   deque of the last F = 2 * num_motions + 1 kept images -> md_window_push (a ring of
                                        pyramids on the device: every image is converted and its pyramid built once)
  callback body      node.cpp:262-395   trajectories -> fitSubspace -> clusterEuclidean -> showBoundingBoxes
                                        -> md_window_detect (one call)
  log_contours_      node.cpp:419-430   ml_.writeBoundingBox(rect, global_frame_count_, i): "frame, id, x0, y0, x1, y1"
                                        (common/src/motion_logger.cpp:43-47)

The compute is libmotion_b200.so only (capi.Context); nothing here touches oracle/.
"""
from . import capi


class LiveNode:
    def __init__(self, width, height, pixel_step=10, num_motions=2, skip_frames=1, sigma=0.5, distance_threshold=50.0,
                 seed=1, device=0, min_vector_size=1.0):
        self.num_motions = num_motions
        self.trajectory_size = 2 * num_motions + 1                  # node.cpp:240
        self.skip_frames = max(1, int(skip_frames))
        self.sigma, self.distance_threshold, self.seed = sigma, distance_threshold, seed
        self.ctx = capi.Context(width=width, height=height, max_batch=self.trajectory_size - 1, pixel_step=pixel_step,
                                min_vector_size=min_vector_size, device=device)
        self.global_frame_count = 0                                 # node.cpp:247,409
        self.frame_number = 0                                       # node.cpp:314,397
        self.callbacks = 0
        self.log_rows = []

    def close(self):
        self.ctx.close()

    def on_image(self, image):
        """One imageCallback.  image: [h][w] gray or [h][w][3] rgb8.  Returns None while the image is skipped or the window
        is still filling, else the result dict of Context.window_detect plus 'rectangles'."""
        if self.global_frame_count % self.skip_frames != 0:         # node.cpp:247
            self.global_frame_count += 1
            return None
        fill = self.ctx.window_push(image)
        if fill < self.trajectory_size:                             # image_received_ stays false (node.cpp:248-255)
            self.global_frame_count += 1
            return None
        # the reference's rand() stream runs on across callbacks; here callback k draws with seed + k
        res = self.ctx.window_detect(num_motions=self.num_motions, sigma=self.sigma, distance_threshold=self.distance_threshold,
                                     seed=self.seed + self.callbacks)
        self.callbacks += 1
        self.frame_number += 1
        res["rectangles"] = res["boxes"]
        for i, b in enumerate(res["boxes"]):                        # motion_logger.cpp:43-47
            self.log_rows.append("%d, %d, %d, %d, %d, %d" % (self.global_frame_count, i, b[0], b[1], b[2], b[3]))
        res["frame"] = self.global_frame_count
        self.global_frame_count += 1
        return res
