"""CPU restatement (numpy) of the colored_scan debug cloud — TEST INFRASTRUCTURE, never the product.

Follows the node's use of ColorPointsByLabel (extraction/app/feature_extraction.cpp:153 with
extraction/include/lidar_feature_extraction/color_points.hpp:46-74 and LabelToColor,
extraction/src/color_points.cpp:39-68): for every ring that did not throw, in ring order, one pcl::PointXYZRGB per
point of the ring-sorted order with x,y,z of the point and the colour of its final label.

Pinned: the colour table and the x,y,z,r,g,b values by the reference's own known-answer test
(extraction/test/test_color_points.cpp:40-78, replayed in tests/test_color_oracle.py). The BYTE layout of
pcl::PointXYZRGB (x,y,z,1.0f at 0..15; b,g,r,a=255 at 16..19; 12 zero bytes) and the PointCloud2 fields
pcl::toROSMsg derives from it are PCL's (third party, absent from /root/reference, PCL 1.12 on the CI image):
parity of those bytes is UNPINNED by any reference test.
"""
from __future__ import annotations

import numpy as np

# LabelToColor, color_points.cpp:39-68 (r, g, b) by PointLabel value
LABEL_RGB = np.array([(255, 255, 255), (255, 0, 0), (255, 63, 0), (255, 0, 0), (255, 63, 0), (127, 127, 127),
                      (255, 0, 255), (0, 255, 0)], np.uint8)


def color_points_by_label(xyz: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """ColorPointsByLabel + MakeXYZRGB, color_points.hpp:46-74 -> [n, 32] uint8 pcl::PointXYZRGB records."""
    n = len(labels)
    if np.any(labels > 7):
        raise ValueError("Invalid label")   # ThrowIfInvalidLabelDetected, color_points.cpp:33-37
    out = np.zeros((n, 32), np.uint8)
    out[:, 0:12] = np.ascontiguousarray(xyz, np.float32).view(np.uint8).reshape(n, 12)
    out[:, 12:16] = np.frombuffer(np.float32(1.0).tobytes(), np.uint8)
    rgb = LABEL_RGB[labels]
    out[:, 16], out[:, 17], out[:, 18], out[:, 19] = rgb[:, 2], rgb[:, 1], rgb[:, 0], 255
    return out


def colored_scan(x, y, z, ref) -> np.ndarray:
    """ref: oracle ScanResult (ring-sorted arrays over the kept rings). Skipped rings contribute nothing
    (the try block of feature_extraction.cpp:126-156 is left before line 153)."""
    parts, pos = [], 0
    for n, skipped in zip(ref.ring_sizes, ref.ring_skipped):
        n = int(n)
        if not skipped:
            src = ref.sorted_src[pos:pos + n]
            parts.append(color_points_by_label(np.stack([x[src], y[src], z[src]], axis=1), ref.labels[pos:pos + n]))
        pos += n
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, 32), np.uint8)
