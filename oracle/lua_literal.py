"""Statement-by-statement restatements of the reference's criterion files -- TEST INFRASTRUCTURE, NOT PRODUCT.

The closed forms in oracle/b2f_oracle.py are what the CUDA kernels are compared with; nobody can run them against
Torch7 here ("parity unpinned").  This module is the second, independent reading of the same Lua:
every statement of `updateOutput` / `updateGradInput` of

    criterions/OBCCriterion.lua:36-240          criterions/OBGCCriterion.lua:39-300
    criterions/SecondOrderSmoothnessCriterion.lua:28-104
    criterions/ConstVelCriterion.lua:29-74      criterions/OcclusionPriorCriterion.lua:28-73
    criterions/penalty/{quadratic,L1,Lorentzian}_function.lua
    models/CostVolMulti.lua:49-181

is written as the same sequence of tensor-method calls on the Torch7 tensor model of oracle/th7.py (views that share
storage, in-place methods, overwrite-`add`, keep-dim reductions, Byte masks), in the order and with the temporaries
the Lua has -- no algebra, no fusion, no re-derivation.  tests/test_lua_literal.py cross-checks the closed forms
against it.  A comment `-- :NN` gives the Lua line a statement restates.
"""
from __future__ import annotations

from . import th7 as torch
from .th7 import ALL, LuaTable


# ---- criterions/penalty/*.lua --------------------------------------------------------------------------------

class QuadraticPenalty:
    def apply(self, x):
        return torch.pow_(x, 2)                                               # quadratic_function.lua:16

    def der(self, x):
        return torch.mul(x, 2)                                                # :20


class L1Penalty:
    def __init__(self, alpha=None):
        self.eps = 0.001 * 0.001                                              # L1_function.lua:16
        self.alpha = 0.5 or alpha                                             # :17 (Lua: `0.5 or alpha` is 0.5)

    def apply(self, x):
        return torch.pow_(x, 2).add(self.eps).pow(self.alpha)                 # :21

    def der(self, x):
        return torch.mul(x, 2 * self.alpha).cdiv(torch.pow_(x, 2).add(self.eps).pow(1 - self.alpha))   # :25


class LorentzianPenalty:
    def __init__(self):
        self.eps = 0.05                                                       # Lorentzian_function.lua:16
        self.eps_sq = self.eps * self.eps                                     # :17

    def set_eps(self, eps):
        self.eps = eps                                                        # :21
        self.eps_sq = self.eps * self.eps                                     # :22

    def apply(self, x):
        return torch.log(1 + 0.5 * torch.div(torch.pow_(x, 2), self.eps_sq))  # :26

    def der(self, x):
        return torch.cdiv(torch.mul(x, 2), torch.add(torch.pow_(x, 2), 2 * self.eps_sq))   # :30


# ---- criterions/OBCCriterion.lua ------------------------------------------------------------------------------

class OBCCriterion:
    def __init__(self):                                                       # :26-35
        self.sizeAverage = True
        self.gradCheck = False
        self.p = QuadraticPenalty()
        self.penalty_out = 1.0
        self.F = 3
        self.pwc_flow_scaling = 1
        self.past_flow = False

    def _mask(self, tcoord, w, h):
        """The five statements both files repeat four times (OBCCriterion.lua:97-101, 158-162, 192-196)."""
        mask = torch.ge(tcoord[ALL, (1,), ALL, ALL], 1)                       # left
        mask.cmul(torch.ge(tcoord[ALL, (2,), ALL, ALL], 1))                   # top
        mask.cmul(torch.le(tcoord[ALL, (1,), ALL, ALL], w))                   # right
        mask.cmul(torch.le(tcoord[ALL, (2,), ALL, ALL], h))                   # bottom
        mask = mask.cuda()
        return mask

    def updateOutput(self, input, target):
        assert len(input) >= 4, "expecting at least four inputs"              # :37
        warp_start = 3                                                        # :40
        if self.past_flow:                                                    # :41
            warp_start = 4
        norm = input[warp_start].size(2) / (input[warp_start].nElement())     # :46
        ref = 0.5 * (self.F - 1)                                              # :48
        b = input[1].size(1)                                                  # :49
        h = input[1].size(3)                                                  # :50
        w = input[1].size(4)                                                  # :51
        self.coord = input[1].clone()                                         # :54
        self.coord[ALL, (1,), ALL, ALL] = torch.range_(1, w).repeatTensor(b, 1, h, 1)                    # :55
        self.coord[ALL, (2,), ALL, ALL] = torch.range_(1, h).repeatTensor(b, 1, w, 1).transpose(3, 4)    # :56
        acc = torch.Tensor_(b, 1, h, w).zero()                                # :59
        occ = input[warp_start - 1]                                           # :65
        for f in range(1, (self.F - 1) + 1):                                  # :68
            img = input[warp_start - 1 + f]                                   # :69
            assert img.nElement() == target.nElement(), "input and target size mismatch"   # :71
            buffer = torch.add(img, -1, target)                               # :74
            tmp = torch.sum_(self.p.apply(buffer), 2)                         # :75
            if f <= ref:                                                      # :79
                if self.past_flow:                                            # :80
                    tcoord = self.coord + (f - ref - 1) * input[2] * self.pwc_flow_scaling   # :81
                else:
                    tcoord = self.coord + (f - ref - 1) * input[1] * self.pwc_flow_scaling   # :83
                tocc = occ[ALL, (2,), ALL, ALL]                               # :86
                tmp.cmul(tocc)                                                # :87
            else:
                tcoord = self.coord + (f - ref) * input[1] * self.pwc_flow_scaling           # :89
                tocc = occ[ALL, (1,), ALL, ALL]                               # :91
                tmp.cmul(tocc)                                                # :92
            if self.gradCheck is False:                                       # :96
                mask = self._mask(tcoord, w, h)                               # :97-101
                tmp.cmul(mask)                                                # :102
                pen = (1 - mask) * self.penalty_out                           # :105
                tmp.add(pen)                                                  # :106
            acc.add(tmp)                                                      # :109
        self.output = acc.sum() / (input[warp_start].size(2) * (self.F - 1))  # :113
        if self.sizeAverage:                                                  # :114
            self.output = norm * self.output                                  # :115
        return self.output

    def updateGradInput(self, input, target):
        assert len(input) >= 4, "expecting at least four inputs"              # :122
        warp_start = 3                                                        # :125
        if self.past_flow:
            warp_start = 4
        norm = input[warp_start].size(2) / (input[warp_start].nElement())     # :130
        gradInput = LuaTable()                                                # :132
        for f in range(1, self.F + 1):                                        # :133
            gradInput.insert(input[warp_start - 2 + f].new())                 # :134
        ref = 0.5 * (self.F - 1)                                              # :137
        b = input[1].size(1)
        h = input[1].size(3)
        w = input[1].size(4)
        occ = input[warp_start - 1]                                           # :142
        gradInput[1].resizeAs(input[warp_start - 1]).fill(0)                  # :144
        for f in range(1, (self.F - 1) + 1):                                  # :147
            img = input[warp_start - 1 + f]                                   # :148
            assert img.nElement() == target.nElement(), "input and target size mismatch"
            buffer = torch.add(img, -1, target)                               # :153
            gradInput[1 + f].resizeAs(buffer)                                 # :156
            gradInput[1 + f].copy(self.p.der(buffer))                         # :157
            buffer = torch.sum_(self.p.apply(buffer), 2)                      # :160
            if f <= ref:                                                      # :161
                if self.gradCheck is False:                                   # :163
                    if self.past_flow:                                        # :166
                        tcoord = self.coord + (f - ref - 1) * input[2] * self.pwc_flow_scaling   # :167
                    else:
                        tcoord = self.coord + (f - ref - 1) * input[1] * self.pwc_flow_scaling   # :169
                    mask = self._mask(tcoord, w, h)                           # :173-177
                    buffer.cmul(mask)                                         # :180
                    pen = (1 - mask) * self.penalty_out                       # :181
                    buffer.add(pen)                                           # :182
                    mask = torch.repeatTensor(mask, 1, input[warp_start].size(2), 1, 1)          # :185
                    gradInput[1 + f].cmul(mask)                               # :186
                gradInput[1][ALL, (2,), ALL, ALL].add(buffer)                 # :190
                tocc = occ[ALL, (2,), ALL, ALL].repeatTensor(1, input[warp_start].size(2), 1, 1)  # :193
                gradInput[1 + f].cmul(tocc)                                   # :194
            else:
                if self.gradCheck is False:                                   # :197
                    tcoord = self.coord + (f - ref) * input[1] * self.pwc_flow_scaling           # :199
                    mask = self._mask(tcoord, w, h)                           # :202-206
                    buffer.cmul(mask)                                         # :209
                    pen = (1 - mask) * self.penalty_out                       # :210
                    buffer.add(pen)                                           # :211
                    mask = torch.repeatTensor(mask, 1, input[warp_start].size(2), 1, 1)          # :214
                    gradInput[1 + f].cmul(mask)                               # :215
                gradInput[1][ALL, (1,), ALL, ALL].add(buffer)                 # :219
                tocc = occ[ALL, (1,), ALL, ALL].repeatTensor(1, input[warp_start].size(2), 1, 1)  # :222
                gradInput[1 + f].cmul(tocc)                                   # :223
            gradInput[1 + f].mul(1 / (input[warp_start].size(2) * (self.F - 1)))                 # :227
            if self.sizeAverage:                                              # :228
                gradInput[1 + f].mul(norm)                                    # :229
        gradInput[1].mul(1 / (input[warp_start].size(2) * (self.F - 1)))      # :234
        if self.sizeAverage:                                                  # :235
            gradInput[1].mul(norm)                                            # :236
        return gradInput


# ---- criterions/OBGCCriterion.lua -----------------------------------------------------------------------------

class OBGCCriterion(OBCCriterion):
    def __init__(self):                                                       # :25-37
        super().__init__()
        self.alpha = 1.0
        self.beta = 1.0
        self.gamma = 1.0

    def _gradient_buffers(self, target):
        """:56-68 and :172-184: four zeroed buffers, the target's forward differences written into the sub-regions."""
        target_gy = torch.Tensor_(target.size()).zero()
        target_gx = torch.Tensor_(target.size()).zero()
        img_gy = torch.Tensor_(target.size()).zero()
        img_gx = torch.Tensor_(target.size()).zero()
        H, W = target.size(3), target.size(4)
        target_gy[ALL, ALL, (1, H - 1), ALL].add(target[ALL, ALL, (2, H), ALL], -1, target[ALL, ALL, (1, H - 1), ALL])
        target_gx[ALL, ALL, ALL, (1, W - 1)].add(target[ALL, ALL, ALL, (2, W)], -1, target[ALL, ALL, ALL, (1, W - 1)])
        return target_gy, target_gx, img_gy, img_gx

    def updateOutput(self, input, target):
        assert len(input) >= 4, "expecting at least four inputs"              # :40
        warp_start = 3
        if self.past_flow:
            warp_start = 4
        norm = input[warp_start].size(2) / (input[warp_start].nElement())     # :49
        ref = 0.5 * (self.F - 1)
        b = input[1].size(1)
        h = input[1].size(3)
        w = input[1].size(4)
        target_gy, target_gx, img_gy, img_gx = self._gradient_buffers(target)  # :56-68
        self.coord = input[1].clone()                                         # :71
        self.coord[ALL, (1,), ALL, ALL] = torch.range_(1, w).repeatTensor(b, 1, h, 1)                    # :72
        self.coord[ALL, (2,), ALL, ALL] = torch.range_(1, h).repeatTensor(b, 1, w, 1).transpose(3, 4)    # :73
        acc = torch.Tensor_(b, 1, h, w).zero()                                # :76
        occ = input[warp_start - 1]                                           # :82
        for f in range(1, (self.F - 1) + 1):                                  # :85
            img = input[warp_start - 1 + f]                                   # :86
            assert img.nElement() == target.nElement(), "input and target size mismatch"
            H, W = img.size(3), img.size(4)
            img_gy[ALL, ALL, (1, H - 1), ALL].add(img[ALL, ALL, (2, H), ALL], -1, img[ALL, ALL, (1, H - 1), ALL])   # :91
            img_gx[ALL, ALL, ALL, (1, W - 1)].add(img[ALL, ALL, ALL, (2, W)], -1, img[ALL, ALL, ALL, (1, W - 1)])   # :92
            buffer = torch.add(img, -1, target)                               # :96
            tmp = torch.sum_(self.p.apply(buffer), 2)                         # :97  (alpha is not applied here: Q5)
            buffer_gx = torch.add(img_gx, -1, target_gx)                      # :100
            tmp.add(torch.sum_(self.p.apply(buffer_gx), 2).mul(self.beta))    # :101
            buffer_gy = torch.add(img_gy, -1, target_gy)                      # :104
            tmp.add(torch.sum_(self.p.apply(buffer_gy), 2).mul(self.gamma))   # :105
            if f <= ref:                                                      # :109
                if self.past_flow:
                    tcoord = self.coord + (f - ref - 1) * input[2] * self.pwc_flow_scaling   # :111
                else:
                    tcoord = self.coord + (f - ref - 1) * input[1] * self.pwc_flow_scaling   # :113
                tocc = occ[ALL, (2,), ALL, ALL]                               # :116
                tmp.cmul(tocc)
            else:
                tcoord = self.coord + (f - ref) * input[1] * self.pwc_flow_scaling           # :119
                tocc = occ[ALL, (1,), ALL, ALL]                               # :121
                tmp.cmul(tocc)
            if self.gradCheck is False:                                       # :126
                mask = self._mask(tcoord, w, h)                               # :127-131
                tmp.cmul(mask)                                                # :132
                pen = (1 - mask) * self.penalty_out                           # :135
                tmp.add(pen)                                                  # :136
            acc.add(tmp)                                                      # :139
        self.output = acc.sum() / (input[warp_start].size(2) * (self.F - 1))  # :143
        if self.sizeAverage:
            self.output = norm * self.output                                  # :145
        return self.output

    def updateGradInput(self, input, target):
        assert len(input) >= 4, "expecting at least four inputs"              # :152
        ref = 0.5 * (self.F - 1)
        b = input[1].size(1)
        h = input[1].size(3)
        w = input[1].size(4)
        warp_start = 3
        if self.past_flow:
            warp_start = 4
        norm = input[warp_start].size(2) / (input[warp_start].nElement())     # :165
        gradInput = LuaTable()
        for f in range(1, self.F + 1):                                        # :168
            gradInput.insert(input[warp_start - 2 + f].new())
        target_gy, target_gx, img_gy, img_gx = self._gradient_buffers(target)  # :172-184
        occ = input[warp_start - 1]                                           # :186
        gradInput[1].resizeAs(input[warp_start - 1]).fill(0)                  # :187
        for f in range(1, (self.F - 1) + 1):                                  # :190
            img = input[warp_start - 1 + f]                                   # :191
            H, W = img.size(3), img.size(4)
            img_gy[ALL, ALL, (1, H - 1), ALL].add(img[ALL, ALL, (2, H), ALL], -1, img[ALL, ALL, (1, H - 1), ALL])   # :194
            img_gx[ALL, ALL, ALL, (1, W - 1)].add(img[ALL, ALL, ALL, (2, W)], -1, img[ALL, ALL, ALL, (1, W - 1)])   # :195
            assert img.nElement() == target.nElement(), "input and target size mismatch"
            buffer = torch.add(img, -1, target)                               # :200
            gradInput[1 + f].resizeAs(buffer)                                 # :201
            gradInput[1 + f].copy(self.p.der(buffer).mul(self.alpha))         # :202
            buffer_gy = torch.add(img_gy, -1, target_gy)                      # :205
            gradInput[1 + f].add(-1, self.p.der(buffer_gy).mul(self.gamma))   # :206
            gradInput[1 + f][ALL, ALL, (2, H), ALL].add(self.p.der(buffer_gy[ALL, ALL, (1, H - 1), ALL]).mul(self.gamma))   # :207
            buffer_gx = torch.add(img_gx, -1, target_gx)                      # :210
            gradInput[1 + f].add(-1, self.p.der(buffer_gx).mul(self.beta))    # :211
            gradInput[1 + f][ALL, ALL, ALL, (2, W)].add(self.p.der(buffer_gx[ALL, ALL, ALL, (1, W - 1)]).mul(self.beta))    # :212
            buffer = torch.sum_(self.p.apply(buffer), 2).mul(self.alpha)      # :215
            buffer.add(-1, torch.sum_(self.p.apply(buffer_gy), 2).mul(self.gamma))                                          # :216
            buffer[ALL, ALL, (2, H), ALL].add(torch.sum_(self.p.apply(buffer_gy[ALL, ALL, (1, H - 1), ALL]), 2).mul(self.gamma))   # :217
            buffer.add(-1, torch.sum_(self.p.apply(buffer_gx), 2).mul(self.beta))                                           # :218
            buffer[ALL, ALL, ALL, (2, W)].add(torch.sum_(self.p.apply(buffer_gx[ALL, ALL, ALL, (1, W - 1)]), 2).mul(self.beta))    # :219
            if f <= ref:                                                      # :221
                if self.gradCheck is False:                                   # :223
                    if self.past_flow:
                        tcoord = self.coord + (f - ref - 1) * input[2] * self.pwc_flow_scaling   # :227
                    else:
                        tcoord = self.coord + (f - ref - 1) * input[1] * self.pwc_flow_scaling   # :229
                    mask = self._mask(tcoord, w, h)                           # :233-237
                    buffer.cmul(mask)                                         # :240
                    pen = (1 - mask) * self.penalty_out                       # :241
                    buffer.add(pen)                                           # :242
                    mask = torch.repeatTensor(mask, 1, input[warp_start].size(2), 1, 1)          # :245
                    gradInput[1 + f].cmul(mask)                               # :246
                gradInput[1][ALL, (2,), ALL, ALL].add(buffer)                 # :250
                tocc = occ[ALL, (2,), ALL, ALL].repeatTensor(1, input[warp_start].size(2), 1, 1)  # :253
                gradInput[1 + f].cmul(tocc)                                   # :254
            else:
                if self.gradCheck is False:                                   # :257
                    tcoord = self.coord + (f - ref) * input[1] * self.pwc_flow_scaling           # :259
                    mask = self._mask(tcoord, w, h)                           # :262-266
                    buffer.cmul(mask)                                         # :269
                    pen = (1 - mask) * self.penalty_out                       # :270
                    buffer.add(pen)                                           # :271
                    mask = torch.repeatTensor(mask, 1, input[warp_start].size(2), 1, 1)          # :274
                    gradInput[1 + f].cmul(mask)                               # :275
                gradInput[1][ALL, (1,), ALL, ALL].add(buffer)                 # :279
                tocc = occ[ALL, (1,), ALL, ALL].repeatTensor(1, input[warp_start].size(2), 1, 1)  # :282
                gradInput[1 + f].cmul(tocc)                                   # :283
            gradInput[1 + f].mul(1 / (input[warp_start].size(2) * (self.F - 1)))                 # :287
            if self.sizeAverage:
                gradInput[1 + f].mul(norm)                                    # :289
        gradInput[1].mul(1 / (input[warp_start].size(2) * (self.F - 1)))      # :294
        if self.sizeAverage:
            gradInput[1].mul(norm)                                            # :296
        return gradInput


# ---- criterions/SecondOrderSmoothnessCriterion.lua ---------------------------------------------------------------

class SecondOrderSmoothnessCriterion:
    def __init__(self):                                                       # :20-26
        self.sizeAverage = True
        self.gradCheck = False
        self.p = QuadraticPenalty()
        self.cs = 20
        self.buffer = None

    def updateOutput(self, input, target):
        assert input.size(3) == target.size(3) and input.size(4) == target.size(4), "input and target size mismatch"
        self.buffer = self.buffer or input.new()                              # :32
        buffer = self.buffer                                                  # :34
        norm = 1.0 / input.nElement()                                         # :35
        self.gy = torch.Tensor_(input.size()).zero()                          # :38
        self.gx = torch.Tensor_(input.size()).zero()                          # :39
        H, W = input.size(3), input.size(4)
        self.gy[ALL, ALL, (2, H - 1), ALL].add(2 * input[ALL, ALL, (2, H - 1), ALL], -1, input[ALL, ALL, (1, H - 2), ALL]) \
            .add(-1, input[ALL, ALL, (3, H), ALL])                            # :45
        self.gx[ALL, ALL, ALL, (2, W - 1)].add(2 * input[ALL, ALL, ALL, (2, W - 1)], -1, input[ALL, ALL, ALL, (1, W - 2)]) \
            .add(-1, input[ALL, ALL, ALL, (3, W)])                            # :46
        igy = torch.Tensor_(input.size(1), 1, input.size(3), input.size(4)).zero()   # :49
        igx = torch.Tensor_(input.size(1), 1, input.size(3), input.size(4)).zero()   # :50
        TH, TW = target.size(3), target.size(4)
        igy[ALL, ALL, (2, TH), ALL].add(torch.mean(torch.add(target[ALL, ALL, (2, TH), ALL], -1, target[ALL, ALL, (1, TH - 1), ALL]).abs(), 2))       # :55
        igx[ALL, ALL, ALL, (2, TW)].add(torch.mean(torch.add(target[ALL, ALL, ALL, (2, TW)], -1, target[ALL, ALL, ALL, (1, TW - 1)]).abs(), 2))       # :56
        igy[ALL, ALL, (2, TH - 1), ALL].add(torch.mean(torch.add(target[ALL, ALL, (2, TH - 1), ALL], -1, target[ALL, ALL, (3, TH), ALL]).abs(), 2))   # :57
        igx[ALL, ALL, ALL, (2, TW - 1)].add(torch.mean(torch.add(target[ALL, ALL, ALL, (2, TW - 1)], -1, target[ALL, ALL, ALL, (3, TW)]).abs(), 2))   # :58
        self.wy = torch.expandAs(torch.exp(-self.cs * igy), self.gy)          # :60
        self.wx = torch.expandAs(torch.exp(-self.cs * igx), self.gx)          # :61
        buffer.resizeAs(input)                                                # :64
        buffer.add(self.p.apply(self.gx).cmul(self.wx), self.p.apply(self.gy).cmul(self.wy))   # :65
        buffer = buffer.sum()                                                 # :66
        if self.sizeAverage:                                                  # :68
            self.output = norm * buffer
        else:
            self.output = buffer
        return self.output

    def updateGradInput(self, input, target):
        assert input.size(3) == target.size(3) and input.size(4) == target.size(4), "input and target size mismatch"
        norm = 1. / input.nElement()                                          # :85
        self.gy = self.p.der(self.gy).cmul(self.wy)                           # :87
        self.gx = self.p.der(self.gx).cmul(self.wx)                           # :88
        H, W = input.size(3), input.size(4)
        gradInput = input.new()
        gradInput.resizeAs(input).zero()                                      # :91
        gradInput[ALL, ALL, (2, H - 1), ALL].add(2 * self.gy[ALL, ALL, (2, H - 1), ALL])        # :92
        gradInput[ALL, ALL, ALL, (2, W - 1)].add(2 * self.gx[ALL, ALL, ALL, (2, W - 1)])        # :93
        gradInput[ALL, ALL, (1, H - 2), ALL].add(-1, self.gy[ALL, ALL, (2, H - 1), ALL])        # :94
        gradInput[ALL, ALL, ALL, (1, W - 2)].add(-1, self.gx[ALL, ALL, ALL, (2, W - 1)])        # :95
        gradInput[ALL, ALL, (3, H), ALL].add(-1, self.gy[ALL, ALL, (2, H - 1), ALL])            # :96
        gradInput[ALL, ALL, ALL, (3, W)].add(-1, self.gx[ALL, ALL, ALL, (2, W - 1)])            # :97
        if self.sizeAverage:                                                  # :99
            gradInput.mul(norm)
        return gradInput


# ---- criterions/SmoothnessCriterion.lua -----------------------------------------------------------------------------

class SmoothnessCriterion:
    def __init__(self):                                                       # :20-26
        self.sizeAverage = True
        self.gradCheck = False
        self.p = QuadraticPenalty()
        self.cs = 20
        self.buffer = None

    def updateOutput(self, input, target):
        assert input.size(3) == target.size(3) and input.size(4) == target.size(4), "input and target size mismatch"
        self.buffer = self.buffer or input.new()                              # :32
        buffer = self.buffer
        norm = 1.0 / input.nElement()                                         # :35
        self.gy = torch.Tensor_(input.size()).zero()                          # :38
        self.gx = torch.Tensor_(input.size()).zero()                          # :39
        H, W = input.size(3), input.size(4)
        self.gy[ALL, ALL, (1, H - 1), ALL].add(input[ALL, ALL, (2, H), ALL], -1, input[ALL, ALL, (1, H - 1), ALL])   # :45
        self.gx[ALL, ALL, ALL, (1, W - 1)].add(input[ALL, ALL, ALL, (2, W)], -1, input[ALL, ALL, ALL, (1, W - 1)])   # :46
        igy = torch.Tensor_(input.size()).zero()                              # :49
        igx = torch.Tensor_(input.size()).zero()                              # :50
        TH, TW = target.size(3), target.size(4)
        if target.size(2) == input.size(2):
            igy[ALL, ALL, (1, TH - 1), ALL].add(target[ALL, ALL, (2, TH), ALL], -1, target[ALL, ALL, (1, TH - 1), ALL])   # :55
            igx[ALL, ALL, ALL, (1, TW - 1)].add(target[ALL, ALL, ALL, (2, TW)], -1, target[ALL, ALL, ALL, (1, TW - 1)])   # :56
        else:
            # :55-56 with a 2-channel input and the 3-channel target (both users of this class, model.lua:216,
            # train.lua:458-462): r:add(a, v, b) resizes the narrowed VIEW to a's size, which re-lays it out contiguously
            # over the shared storage.  th7 refuses that on purpose; the two statements are replayed on
            # the storage-level model instead (oracle/b2f_oracle.py: smooth1_weight_inputs_literal)
            from . import b2f_oracle as _o
            ly, lx = _o.smooth1_weight_inputs_literal(input.size(), target.a, target.a.dtype.type)
            igy, igx = torch.Tensor(ly), torch.Tensor(lx)
        self.wy = torch.expandAs(torch.exp(-self.cs * torch.mean(torch.abs_(igy), 2)), self.gy)   # :58
        self.wx = torch.expandAs(torch.exp(-self.cs * torch.mean(torch.abs_(igx), 2)), self.gx)   # :59
        buffer.resizeAs(input)                                                # :62
        buffer.add(self.p.apply(self.gx).cmul(self.wx), self.p.apply(self.gy).cmul(self.wy))      # :63
        buffer = buffer.sum()                                                 # :64
        if self.sizeAverage:                                                  # :66
            self.output = norm * buffer
        else:
            self.output = buffer
        return self.output

    def updateGradInput(self, input, target):
        assert input.size(3) == target.size(3) and input.size(4) == target.size(4), "input and target size mismatch"
        norm = 1. / input.nElement()                                          # :83
        self.gy = self.p.der(self.gy).cmul(self.wy)                           # :85
        self.gx = self.p.der(self.gx).cmul(self.wx)                           # :86
        gys1 = torch.Tensor_(input.size()).zero()                             # :89
        gxs1 = torch.Tensor_(input.size()).zero()                             # :90
        H, W = input.size(3), input.size(4)
        gys1[ALL, ALL, (2, H), ALL].copy(self.gy[ALL, ALL, (1, H - 1), ALL])  # :95
        gxs1[ALL, ALL, ALL, (2, W)].copy(self.gx[ALL, ALL, ALL, (1, W - 1)])  # :96
        if self.sizeAverage:                                                  # :99
            gradInput = norm * (-self.gx + gxs1 - self.gy + gys1)             # :100
        else:
            gradInput = (-self.gx + gxs1 - self.gy + gys1)                    # :102
        return gradInput


# ---- criterions/ConstVelCriterion.lua ---------------------------------------------------------------------------

class ConstVelCriterion:
    eps = 1e-12                                                               # :27

    def __init__(self):                                                       # :21-25
        self.sizeAverage = True
        self.gradCheck = False

    def updateOutput(self, input):
        assert input[1].nElement() == input[2].nElement(), "input and target size mismatch"   # :30
        norm = 1.0 / input[1].nElement()                                      # :33
        self.output = torch.add(input[1], -1, input[2]).pow(2)                # :36
        self.output = torch.sum_(self.output, 2).sqrt()                       # :37
        self.output = self.output.sum()                                       # :38
        if self.sizeAverage:                                                  # :41
            self.output = norm * self.output
        return self.output

    def updateGradInput(self, input):
        assert input[1].nElement() == input[2].nElement(), "input and target size mismatch"
        npixels = input[1].nElement() / input[1].size(2)                      # :56
        buffer = torch.add(input[1], -1, input[2]).pow(2)                     # :59
        loss = torch.sum_(buffer, 2).sqrt().add(self.eps)                     # :60
        loss = loss.repeatTensor(1, input[1].size(2), 1, 1)                   # :61
        gradInput = LuaTable()
        gradInput[1] = torch.add(input[1], -1, input[2]).cdiv(loss)           # :64
        gradInput[2] = torch.add(input[2], -1, input[1]).cdiv(loss)           # :65
        if self.sizeAverage:                                                  # :68
            gradInput[1] = gradInput[1] / npixels
            gradInput[2] = gradInput[2] / npixels
        return gradInput


# ---- criterions/OcclusionPriorCriterion.lua -----------------------------------------------------------------------

class OcclusionPriorCriterion:
    def __init__(self):                                                       # :22-26
        self.sizeAverage = True
        self.penalty = 1

    def updateOutput(self, input, target):
        assert input.size(3) == target.size(3) and input.size(4) == target.size(4), "input and target size mismatch"
        norm = input.size(2) / input.nElement()                               # :32
        output = input.new()                                                  # :34
        output.resize(input.size(1), 1, input.size(3), input.size(4))         # :35
        if input.size(2) == 3:                                                # :36
            output[ALL, (1,), ALL, ALL] = (1 - input[ALL, (2,), ALL, ALL]).cmul(
                torch.add(input[ALL, (1,), ALL, ALL], input[ALL, (3,), ALL, ALL])) * self.penalty * 0.05   # :37
        else:
            output[ALL, (1,), ALL, ALL] = (1 - torch.cmul(input[ALL, (1,), ALL, ALL], input[ALL, (2,), ALL, ALL])) * self.penalty   # :39
        if self.sizeAverage:                                                  # :42
            output = norm * output.sum()
        else:
            output = output.sum()
        return output

    def updateGradInput(self, input, target):
        assert input.size(3) == target.size(3) and input.size(4) == target.size(4), "input and target size mismatch"
        norm = input.size(2) / input.nElement()                               # :55
        gradInput = input.clone()                                             # :57
        if input.size(2) == 3:                                                # :59
            gradInput[ALL, (1,), ALL, ALL] = (1 - input[ALL, (2,), ALL, ALL]) * self.penalty * 0.05                          # :60
            gradInput[ALL, (2,), ALL, ALL] = -(input[ALL, (1,), ALL, ALL] + input[ALL, (3,), ALL, ALL]) * self.penalty * 0.05   # :61
            gradInput[ALL, (3,), ALL, ALL] = (1 - input[ALL, (2,), ALL, ALL]) * self.penalty * 0.05                          # :62
        else:
            gradInput[ALL, (1,), ALL, ALL] = (1 - input[ALL, (2,), ALL, ALL]) * self.penalty   # :64
            gradInput[ALL, (2,), ALL, ALL] = (1 - input[ALL, (1,), ALL, ALL]) * self.penalty   # :65
        if self.sizeAverage:                                                  # :68
            gradInput.mul(norm)
        return gradInput


# ---- models/CostVolMulti.lua --------------------------------------------------------------------------------------

class CostVolMulti:
    def __init__(self, win=None, fwd=None):                                   # :23-33
        self.win = win or 9
        self.fwd = True if fwd is None else fwd
        self.output = torch.Tensor_(0)
        self.gradInput = LuaTable()

    def _ranges(self, q_x, q_y, w, h):
        """:77-88, the four Lua range tables (1-based, inclusive)"""
        qx = (1 + q_x, w)
        px = (1, w - q_x)
        if q_x < 0:
            qx = (1, w + q_x)
            px = (1 - q_x, w)
        qy = (1 + q_y, h)
        py = (1, h - q_y)
        if q_y < 0:
            qy = (1, h + q_y)
            py = (1 - q_y, h)
        return qx, px, qy, py

    def updateOutput(self, input):
        frames = len(input)                                                   # :50
        for f in range(2, frames + 1):
            assert input[f].nElement() == input[f - 1].nElement(), "input sizes mismatch"   # :52
        ref = input[1]                                                        # :55
        N, h, w = ref.size(2), ref.size(3), ref.size(4)                       # :56
        n = int(0.5 * (self.win - 1))                                         # :57
        self.output.resize(ref.size(1), self.win * self.win, h, w).zero()     # :59
        for f in range(2, frames + 1):                                        # :62
            frame = input[f]                                                  # :63
            i = 1                                                             # :65
            for q_x_ in range(-n, n + 1):                                     # :66
                for q_y_ in range(-n, n + 1):                                 # :67
                    q_x = q_x_ * (f - 1)                                      # :68
                    q_y = q_y_ * (f - 1)                                      # :69
                    if self.fwd is False:                                     # :71
                        q_x = q_x * -1
                        q_y = q_y * -1
                    qx, px, qy, py = self._ranges(q_x, q_y, w, h)             # :76-87
                    cost = torch.cmul(ref[ALL, ALL, qy, qx], frame[ALL, ALL, py, px])   # :89
                    self.output[ALL, i, qy, qx].add(cost.sum(2))              # :90
                    i = i + 1                                                 # :92
        self.output.div(N * (frames - 1))                                     # :100
        return self.output

    def updateGradInput(self, input, gradOutput):
        frames = len(input)                                                   # :112
        ref = input[1]                                                        # :114
        bs, N, h, w = ref.size(1), ref.size(2), ref.size(3), ref.size(4)      # :115
        n = int(0.5 * (self.win - 1))                                         # :116
        if len(self.gradInput) != frames:                                     # :118
            self.gradInput = LuaTable([input[f].new() for f in range(1, frames + 1)])
        for f in range(1, frames + 1):                                        # :124
            self.gradInput[f].resizeAs(input[f]).zero()
        gradInputRef = self.gradInput[1]                                      # :128
        for f in range(2, frames + 1):                                        # :131
            frame = input[f]
            gradInputFrame = self.gradInput[f]                                # :133
            i = 1
            for q_x_ in range(-n, n + 1):                                     # :136
                for q_y_ in range(-n, n + 1):                                 # :137
                    q_x = q_x_ * (f - 1)
                    q_y = q_y_ * (f - 1)
                    if self.fwd is False:                                     # :141
                        q_x = q_x * -1
                        q_y = q_y * -1
                    qx, px, qy, py = self._ranges(q_x, q_y, w, h)             # :146-157
                    ny = qy[1] - qy[0] + 1                                    # :159
                    nx = qx[1] - qx[0] + 1                                    # :160
                    go = gradOutput[ALL, i, qy, qx].clone().view(bs, 1, ny, nx)   # :161
                    go = torch.repeatTensor(go, 1, N, 1, 1)                   # :162
                    gradInputRef[ALL, ALL, qy, qx].add(torch.cmul(go, frame[ALL, ALL, py, px]))     # :164
                    gradInputFrame[ALL, ALL, py, px].add(torch.cmul(go, ref[ALL, ALL, qy, qx]))     # :165
                    i = i + 1                                                 # :168
        for f in range(1, frames + 1):                                        # :176
            self.gradInput[f].div(N * (frames - 1))
        return self.gradInput
