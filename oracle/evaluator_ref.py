"""ORACLE — test infrastructure only.  Restatement of the reference's evaluator inference
(engine/evaluator.py:433-558: sliding_eval_rgbX, scale_process_rgbX, val_func_process_rgbX, process_image_rgbX) and of
TrainPre (dataloader/dataloader.py:8-50) with the model forward passed in as `val_func` (numpy in, numpy out), using cv2 and
numpy exactly where the reference does (cv2 is a third-party dependency of the reference, present in this image).
The reference's quirks are kept on purpose: window x-extents use crop_size[0] and y-extents crop_size[1] (:472-477)."""
import cv2
import numpy as np


def normalize(img, mean, std):
    """utils/transforms.py:182-187"""
    img = img.astype(np.float64) / 255.0
    return (img - mean) / std


def pad_image_to_shape(img, shape, value=0):
    """utils/transforms.py:61-75 (cv2.BORDER_CONSTANT)"""
    margin = np.zeros(4, np.uint32)
    pad_height = shape[0] - img.shape[0] if shape[0] - img.shape[0] > 0 else 0
    pad_width = shape[1] - img.shape[1] if shape[1] - img.shape[1] > 0 else 0
    margin[0] = pad_height // 2
    margin[1] = pad_height // 2 + pad_height % 2
    margin[2] = pad_width // 2
    margin[3] = pad_width // 2 + pad_width % 2
    img = cv2.copyMakeBorder(img, int(margin[0]), int(margin[1]), int(margin[2]), int(margin[3]), cv2.BORDER_CONSTANT, value=value)
    return img, margin


def process_image_rgbX(img, modal_x, crop_size, mean, std):
    """evaluator.py:523-558 (3-channel modal_x)"""
    p_img = normalize(img, mean, std)
    p_x = normalize(modal_x, mean, std)
    p_img, margin = pad_image_to_shape(p_img, crop_size, 0)
    p_x, _ = pad_image_to_shape(p_x, crop_size, 0)
    return p_img.transpose(2, 0, 1), p_x.transpose(2, 0, 1), margin


def val_func_process_rgbX(val_func, input_data, input_modal_x, is_flip):
    """evaluator.py:501-522: val_func(rgb (1,3,H,W) f32, x (1,3,H,W) f32) -> (1,C,H,W) f32; score = exp(s (+ flip))"""
    a = np.ascontiguousarray(input_data[None], dtype=np.float32)
    b = np.ascontiguousarray(input_modal_x[None], dtype=np.float32)
    score = val_func(a, b)[0]
    if is_flip:
        sf = val_func(np.ascontiguousarray(a[..., ::-1]), np.ascontiguousarray(b[..., ::-1]))[0]
        score = score + sf[..., ::-1]
    return np.exp(score.astype(np.float32))


def scale_process_rgbX(val_func, img, modal_x, ori_shape, crop_size, stride_rate, mean, std, is_flip, class_num):
    """evaluator.py:454-499"""
    new_rows, new_cols, _ = img.shape
    if new_cols <= crop_size[1] or new_rows <= crop_size[0]:
        input_data, input_modal_x, margin = process_image_rgbX(img, modal_x, crop_size, mean, std)
        score = val_func_process_rgbX(val_func, input_data, input_modal_x, is_flip)
        score = score[:, margin[0]:(score.shape[1] - margin[1]), margin[2]:(score.shape[2] - margin[3])]
    else:
        stride = (int(np.ceil(crop_size[0] * stride_rate)), int(np.ceil(crop_size[1] * stride_rate)))
        img_pad, margin = pad_image_to_shape(img, crop_size, 0)
        modal_x_pad, margin = pad_image_to_shape(modal_x, crop_size, 0)
        pad_rows, pad_cols = img_pad.shape[0], img_pad.shape[1]
        r_grid = int(np.ceil((pad_rows - crop_size[0]) / stride[0])) + 1
        c_grid = int(np.ceil((pad_cols - crop_size[1]) / stride[1])) + 1
        data_scale = np.zeros((class_num, pad_rows, pad_cols), np.float32)
        for grid_yidx in range(r_grid):
            for grid_xidx in range(c_grid):
                s_x = grid_xidx * stride[0]
                s_y = grid_yidx * stride[1]
                e_x = min(s_x + crop_size[0], pad_cols)
                e_y = min(s_y + crop_size[1], pad_rows)
                s_x = e_x - crop_size[0]
                s_y = e_y - crop_size[1]
                img_sub = img_pad[s_y:e_y, s_x:e_x, :]
                modal_x_sub = modal_x_pad[s_y:e_y, s_x:e_x, :]
                input_data, input_modal_x, tmargin = process_image_rgbX(img_sub, modal_x_sub, crop_size, mean, std)
                temp_score = val_func_process_rgbX(val_func, input_data, input_modal_x, is_flip)
                temp_score = temp_score[:, tmargin[0]:(temp_score.shape[1] - tmargin[1]), tmargin[2]:(temp_score.shape[2] - tmargin[3])]
                data_scale[:, s_y:e_y, s_x:e_x] += temp_score
        score = data_scale
        score = score[:, margin[0]:(score.shape[1] - margin[1]), margin[2]:(score.shape[2] - margin[3])]
    score = score.transpose(1, 2, 0)
    return cv2.resize(np.ascontiguousarray(score), (ori_shape[1], ori_shape[0]), interpolation=cv2.INTER_LINEAR)


def sliding_eval_rgbX(val_func, img, modal_x, crop_size, stride_rate, multi_scales, is_flip, class_num, mean, std):
    """evaluator.py:433-452 -> pred (H, W) int64"""
    ori_rows, ori_cols, _ = img.shape
    processed_pred = np.zeros((ori_rows, ori_cols, class_num))
    for s in multi_scales:
        img_scale = cv2.resize(img, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR)
        modal_x_scale = cv2.resize(modal_x, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR)
        processed_pred += scale_process_rgbX(val_func, img_scale, modal_x_scale, (ori_rows, ori_cols), crop_size, stride_rate, mean, std,
                                             is_flip, class_num)
    return processed_pred.argmax(2)


def train_pre(rgb, gt, modal_x, mirror, scale, crop_pos, crop_size, mean, std):
    """dataloader.py:8-50 with the random draws (mirror, scale or None, crop_pos) passed in; random_crop_pad_to_shape
    (utils/transforms.py:24-42): crop at crop_pos, pad to the crop size with 0 (images) / 255 (labels)."""
    if mirror:
        rgb, gt, modal_x = cv2.flip(rgb, 1), cv2.flip(gt, 1), cv2.flip(modal_x, 1)
    if scale is not None:
        sh, sw = int(rgb.shape[0] * scale), int(rgb.shape[1] * scale)
        rgb = cv2.resize(rgb, (sw, sh), interpolation=cv2.INTER_LINEAR)
        gt = cv2.resize(gt, (sw, sh), interpolation=cv2.INTER_NEAREST)
        modal_x = cv2.resize(modal_x, (sw, sh), interpolation=cv2.INTER_LINEAR)
    rgb, modal_x = normalize(rgb, mean, std), normalize(modal_x, mean, std)

    def crop_pad(img, value):   # utils/transforms.py:24-42: crop, then pad_image_to_shape (centred margins)
        y0, x0 = crop_pos
        c = img[y0:y0 + crop_size[0], x0:x0 + crop_size[1], ...]
        return pad_image_to_shape(c, crop_size, value)[0]
    return crop_pad(rgb, 0).transpose(2, 0, 1), crop_pad(gt, 255), crop_pad(modal_x, 0).transpose(2, 0, 1)
