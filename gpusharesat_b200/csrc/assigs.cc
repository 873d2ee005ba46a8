// assigs.cc -- see assigs.h
#include "assigs.h"
#include <algorithm>
#include <cstring>

namespace gss {

SolverAssigs::SolverAssigs(int varCount) : lastVarVal_(varCount, V_UNDEF), varToUpdatePos_(varCount, -1) {}

void SolverAssigs::setVarCount(int varCount) {
    std::lock_guard<std::mutex> g(lock_);
    if ((int)lastVarVal_.size() < varCount) {
        lastVarVal_.resize(varCount, V_UNDEF);
        varToUpdatePos_.resize(varCount, -1);
    }
}

static inline uint32_t fill(bool b) { return b ? ~0u : 0u; }

void SolverAssigs::setVarLocked(int var, uint8_t val) {
    // Assigs.cu:163-185.  Slots still being written (notCompletedMask) take the new value;
    // frozen slots keep what the device already holds for them, which is lastVarVal.
    bool isSet = val != V_UNDEF, isTrue = val == V_TRUE;
    int pos = varToUpdatePos_[var];
    uint32_t m = notCompletedMask_;
    HostBuf<VarUpdate> &updates_ = updates();
    if (pos == -1 || (int)updates_.size() <= pos || updates_[pos].var != var) {
        VarUpdate vu;
        vu.var = var;
        vu.def = (~m & fill(lastVarVal_[var] != V_UNDEF)) | (m & fill(isSet));
        vu.tru = (~m & fill(lastVarVal_[var] == V_TRUE)) | (m & fill(isTrue));
        varToUpdatePos_[var] = (int)updates_.size();
        if (updates_.size() == updates_.capacity() && allocDevice_ >= 0 && cudaSetDevice(allocDevice_) != cudaSuccess)
            cudaGetLastError(); // (growth page-locks memory: do it in the sharer's context, not in device 0's)
        updates_.push_back(vu);
    } else {
        VarUpdate &vu = updates_[pos];
        vu.def = (vu.def & ~m) | (m & fill(isSet));
        vu.tru = (vu.tru & ~m) | (m & fill(isTrue));
    }
    lastVarVal_[var] = val;
}

int64_t SolverAssigs::assignmentDoneLocked() {
    // Assigs.cu:194-201: freeze the current slot
    int pos = (int)(currentId_ % kSlots);
    GSS_CHECK(firstIdUsed_ + kSlots != currentId_);
    GSS_CHECK(notCompletedMask_ & (1u << pos));
    notCompletedMask_ &= ~(1u << pos);
    return currentId_++;
}

void SolverAssigs::getCurrentAssignment(uint8_t *assig) {
    memcpy(assig, lastVarVal_.data(), lastVarVal_.size());
}

uint32_t SolverAssigs::maskFromTo(int64_t fromId, int64_t toId) {
    uint32_t m = 0;
    for (int64_t i = fromId; i < toId; i++) m |= 1u << (i % kSlots);
    return m;
}

void SolverAssigs::collectLocked(HostBuf<VarUpdate> &out, SolverRunParams &p, AssigIds &ids, bool fullRebuild) {
    HostBuf<VarUpdate> &updates_ = updates();
    const int32_t updStart = (int32_t)out.size();
    if (fullRebuild) {
        ids.start = firstIdUsed_;
        ids.count = (int32_t)(currentId_ - firstIdUsed_);
        // every variable: touched ones carry their update, the others all-slots = lastVarVal
        for (int v = 0; v < (int)lastVarVal_.size(); v++) {
            int pos = varToUpdatePos_[v];
            if (pos != -1 && pos < (int)updates_.size() && updates_[pos].var == v) out.push_back(updates_[pos]);
            else out.push_back(VarUpdate{v, fill(lastVarVal_[v] != V_UNDEF), fill(lastVarVal_[v] == V_TRUE)});
        }
        finishCollectLocked(updStart, (int32_t)out.size() - updStart, p);
    } else {
        collectIntoLocked(out.append(updates_.size()), updStart, p, ids);
    }
}

void SolverAssigs::takeUpdatesLocked(const VarUpdate *&ptr, int32_t updStart, SolverRunParams &p, AssigIds &ids, bool *pinned) {
    ids.start = firstIdUsed_;
    ids.count = (int32_t)(currentId_ - firstIdUsed_);
    HostBuf<VarUpdate> &taken = updates();
    ptr = taken.data();
    *pinned = taken.pinned() || taken.empty();
    finishCollectLocked(updStart, (int32_t)taken.size(), p); // resets the size of `taken`; its records stay where they are
    curUpd_ ^= 1;
    // grow the other buffer to the same capacity now, so that the solver thread rarely allocates
    updates().reserve(taken.capacity());
    updates().clear();
}

void SolverAssigs::collectIntoLocked(VarUpdate *dst, int32_t updStart, SolverRunParams &p, AssigIds &ids) {
    HostBuf<VarUpdate> &updates_ = updates();
    ids.start = firstIdUsed_;
    ids.count = (int32_t)(currentId_ - firstIdUsed_);
    if (!updates_.empty()) memcpy(dst, updates_.data(), updates_.size() * sizeof(VarUpdate));
    finishCollectLocked(updStart, (int32_t)updates_.size(), p);
}

void SolverAssigs::finishCollectLocked(int32_t updStart, int32_t updCount, SolverRunParams &p) {
    p.updStart = updStart;
    p.updCount = updCount;
    updatesSent_ += (int64_t)updates().size();

    // Assigs.cu:255-261: the slot everything collapses to after the run
    int64_t lastIdCopied;
    if (currentId_ == 0) lastIdCopied = 0;
    else if (currentId_ == firstIdUsed_ + kSlots) lastIdCopied = currentId_ - 1;
    else lastIdCopied = currentId_;

    // Assigs.cu:263-284: spread the frozen slots over this solver's aggregate bits
    int bitsUsed = (int)(currentId_ - firstIdUsed_);
    int aggBitsUsed = std::min(endAggBit_ - startAggBit_, bitsUsed);
    p.nGroups = 0;
    p.usedAggBits = 0;
    if (aggBitsUsed != 0) {
        int64_t id = firstIdUsed_;
        int aggBit = startAggBit_;
        int low = bitsUsed / aggBitsUsed, missing = bitsUsed - low * aggBitsUsed;
        for (int i = 0; i < aggBitsUsed; i++) {
            int n = low + (i < missing ? 1 : 0);
            p.groupAggBit[p.nGroups] = 1u << aggBit;
            p.groupSlotMask[p.nGroups] = maskFromTo(id, id + n);
            p.usedAggBits |= 1u << aggBit;
            p.nGroups++;
            aggBit++;
            id += n;
        }
        GSS_CHECK(id == currentId_);
    }
    p.startVals = ~notCompletedMask_;
    p.lastMask = 1u << (lastIdCopied % kSlots);
    p.allAggBits = 0;
    for (int b = startAggBit_; b < endAggBit_; b++) p.allAggBits |= 1u << b;
    p.pad = 0;

    updates().clear();
    firstIdUsed_ = currentId_;
    notCompletedMask_ = ~0u;
}

HostAssigs::HostAssigs() { growSolvers(1); }

void HostAssigs::setVarCount(int varCount) {
    for (auto &s : solvers_) s->setVarCount(varCount);
    varCount_ = std::max(varCount_, varCount);
}

void HostAssigs::growSolvers(int count) {
    int old = (int)solvers_.size();
    if (count > old) {
        solvers_.resize(count);
        for (int i = old; i < count; i++) solvers_[i] = std::make_unique<SolverAssigs>(varCount_);
    }
    count = (int)solvers_.size();
    // Assigs.cu:409-425 hands the 32 aggregate bits of ONE word to all solvers, so solvers
    // 32.. get none and are never checked.  Here every group of 32 solvers has its own
    // aggregate word, partitioned the same way inside the group.
    for (int g0 = 0; g0 < count; g0 += kMaxSolversPerGroup) {
        int n = std::min(kMaxSolversPerGroup, count - g0);
        int low = 32 / n, missing = 32 - low * n, bit = 0;
        for (int i = 0; i < n; i++) {
            int nb = low + (i < missing ? 1 : 0);
            solvers_[g0 + i]->setAggBits(bit, bit + nb);
            bit += nb;
        }
    }
}

} // namespace gss
