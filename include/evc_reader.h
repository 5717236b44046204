/* libevc_reader — native host-side input path of the frame-level YouTube-8M hot path (C ABI, no CUDA).
 *
 * Replaces, for the H-LSTM teacher-student step, what the reference does in TensorFlow's C++ runtime
 * behind code_student_uniform/readers.py:114-246:
 *   tf.TFRecordReader().read                      (readers.py:189-190)  -> TFRecord framing scan (+ optional CRC32C)
 *   tf.parse_single_sequence_example              (readers.py:192-200)  -> SequenceExample wire-format decode
 *   tf.decode_raw(..., tf.uint8) / reshape        (readers.py:166-168)  -> row copy into [max_frames, sum(sizes)]
 *   resize_axis(feature_matrix, 0, max_frames)    (readers.py:173)      -> zero fill past num_frames
 *   tf.sparse_to_dense(labels, validate_indices=False)  (readers.py:203-205) -> dense bool labels
 *   num_frames = min(rows, max_frames)            (readers.py:170), all features must agree (readers.py:218-219)
 * The features stay uint8: utils.Dequantize (utils.py:10-25) runs fused on the GPU (evc_frames_pack_u8).
 *
 * One reader object streams the records of a list of shards in order; every evc_reader_next call decodes the
 * next `batch` records IN PARALLEL (num_threads worker threads) straight into caller-owned buffers (pinned
 * host memory in the product), so no intermediate copy exists.  Not thread-safe per object; use one object
 * per consumer.  All functions return >= 0 on success and < 0 on error (message: evc_reader_last_error()).
 */
#ifndef EVC_READER_H
#define EVC_READER_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct evc_reader evc_reader;

int evc_reader_version(void);
const char* evc_reader_last_error(void);

/* paths: TFRecord shards, read in the given order.  feature_names/feature_sizes: the byte-quantised
 * feature lists to concatenate along the feature axis (readers.py:127-146: e.g. {"rgb","audio"} x {1024,128}).
 * verify_crc != 0: check the masked CRC32C of every record length and payload like tf.TFRecordReader does
 * (a mismatch is an error).  num_threads <= 0: one per hardware thread. */
evc_reader* evc_reader_open(const char* const* paths, int num_paths, const char* const* feature_names,
                            const int* feature_sizes, int num_features, int num_classes, int max_frames,
                            int num_threads, int verify_crc);

/* Decodes up to `batch` videos.  features: uint8 [batch, max_frames, sum(feature_sizes)] (frames past
 * num_frames are zero-filled); labels: uint8 0/1 [batch, num_classes]; num_frames: int32 [batch];
 * ids: char [batch, id_stride], NUL-terminated (truncated to id_stride-1 bytes), may be NULL.
 * Returns the number of videos written (0 = all shards exhausted). */
int evc_reader_next(evc_reader* r, int batch, unsigned char* features, unsigned char* labels, int* num_frames,
                    char* ids, int id_stride);

/* records handed out so far / restart from the first shard */
long long evc_reader_position(const evc_reader* r);
int evc_reader_rewind(evc_reader* r);
void evc_reader_close(evc_reader* r);

/* masked CRC32C as stored in TFRecord files (tensorflow/core/lib/hash/crc32c.h: rotate right 15, + 0xa282ead8) */
unsigned int evc_crc32c_masked(const unsigned char* data, long long n);

#ifdef __cplusplus
}
#endif
#endif
